"""Packet channels between processes (include/cb_chan.h, csrc/exec/cb_chan.c) over POSIX shared-memory arenas: the transport under
the MotionIPCLayer implementation.  Two and three real processes: ordered delivery of packets of every size,
flow control against a slow receiver (the ring never overwrites an unread packet), any-source receives that starve nobody."""
import ctypes as C
import multiprocessing as mp
import os
import random
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Mem(C.Structure):
    _fields_ = [("arg", C.c_void_p), ("put", C.c_void_p), ("get", C.c_void_p)]


def lib():
    L = C.CDLL(os.path.join(ROOT, "cloudberry_b200", "libcbgpu.so"), mode=C.RTLD_GLOBAL)       # noqa: F841 (libcbexec links it)
    E = C.CDLL(os.path.join(ROOT, "cloudberry_b200", "libcbexec.so"))
    E.cb_chan_arena_bytes.restype = C.c_size_t
    E.cb_chan_arena_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    E.cb_chan_shm_create.restype = C.c_void_p
    E.cb_chan_shm_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_size_t]
    E.cb_chan_shm_attach.argtypes = [C.c_void_p, C.POINTER(Mem)]
    E.cb_chan_shm_close.argtypes = [C.c_void_p, C.c_int]
    E.cb_chan_create.restype = C.c_void_p
    E.cb_chan_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Mem)]
    E.cb_chan_destroy.argtypes = [C.c_void_p]
    E.cb_chan_send.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_int]
    E.cb_chan_recv.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int), C.c_int]
    E.cb_chan_max_packet.argtypes = [C.c_void_p]
    return E


SLOTS, SLOT_BYTES = 4, 4096


def payload(src, dst, i):
    rnd = random.Random(src * 1000003 + dst * 10007 + i)
    n = rnd.choice([1, 7, 8, 255, 256, 257, 1000, SLOT_BYTES - 8])
    return bytes(rnd.getrandbits(8) for _ in range(n))


def endpoint(token, rank, world, barrier, conn, npk, slow):
    try:
        E = lib()
        shm = E.cb_chan_shm_create(token.encode(), rank, world, E.cb_chan_arena_bytes(world, SLOTS, SLOT_BYTES))
        assert shm
        mem = Mem()
        t0 = time.time()
        while E.cb_chan_shm_attach(shm, C.byref(mem)) != 0:      # its own rendezvous: peers that are not there yet are waited for
            assert time.time() - t0 < 60
            time.sleep(0.001)
        ch = E.cb_chan_create(rank, world, SLOTS, SLOT_BYTES, C.byref(mem))
        assert ch and E.cb_chan_max_packet(ch) == SLOT_BYTES - 8
        barrier.wait()
        buf = C.create_string_buffer(SLOT_BYTES)
        frm = C.c_int(-1)
        sent = {d: 0 for d in range(world) if d != rank}
        got = {s: 0 for s in range(world) if s != rank}
        crc = 0
        full_seen = 0
        t0 = time.time()
        # everybody sends npk packets to everybody else and receives as many from each, interleaved; a send that finds the
        # ring full (timeout 0) turns to receiving instead of waiting: no deadlock however the processes are scheduled
        while (any(v < npk for v in sent.values()) or any(v < npk for v in got.values())) and time.time() - t0 < 120:
            for d in sent:
                if sent[d] < npk:
                    p = payload(rank, d, sent[d])
                    rc = E.cb_chan_send(ch, d, p, len(p), 0)
                    assert rc in (0, 1), rc
                    if rc == 0:
                        sent[d] += 1
                    else:
                        full_seen += 1
            if slow and rank == 0:
                time.sleep(0.0005)      # a slow receiver: the senders must see full rings, nothing may be lost
            n = E.cb_chan_recv(ch, -1, buf, SLOT_BYTES, C.byref(frm), 0)
            assert n >= 0, n
            if n > 0:
                s = frm.value
                want = payload(s, rank, got[s])
                assert buf.raw[:n] == want, (rank, s, got[s], n, len(want))     # in order, byte for byte
                crc = zlib.crc32(want, crc)
                got[s] += 1
        E.cb_chan_destroy(ch)
        barrier.wait()
        E.cb_chan_shm_close(shm, 1)
        conn.send({"rank": rank, "sent": sent, "got": got, "full_seen": full_seen})
    except BaseException as e:         # noqa: BLE001
        try:
            barrier.abort()
        except Exception:               # noqa: BLE001
            pass
        conn.send("error rank %d: %r" % (rank, e))


def run(world, npk, slow):
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(world)
    token = "cbchan_test_%d_%d" % (os.getpid(), world)
    pipes, procs = [], []
    for r in range(world):
        a, b = ctx.Pipe(duplex=False)
        p = ctx.Process(target=endpoint, args=(token, r, world, barrier, b, npk, slow), daemon=True)
        p.start()
        pipes.append(a)
        procs.append(p)
    res = [a.recv() if a.poll(180) else "error: timed out" for a in pipes]
    for p in procs:
        p.join(10)
    assert not any(isinstance(r, str) for r in res), res
    return res


def test_two_processes_ordered_delivery():
    for r in run(2, 400, False):
        assert all(v == 400 for v in r["sent"].values()) and all(v == 400 for v in r["got"].values())


def test_three_processes_slow_receiver_flow_control():
    res = run(3, 300, True)
    for r in res:
        assert all(v == 300 for v in r["sent"].values()) and all(v == 300 for v in r["got"].values())
    # whoever sent to the slow rank ran into its full ring (4 slots) and lost nothing
    assert sum(r["full_seen"] for r in res if r["rank"] != 0) > 0
