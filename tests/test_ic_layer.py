"""The MotionIPCLayer implementation integration/cbgpu_ic_layer.c, RUN under the reference's own Motion
layer: cdb/motion/cdbmotion.c (SendTuple / SendEndOfStream / RecvTupleFrom / SendStopMessage), tupser.c, tupchunklist.c,
htupfifo.c, heaptuple.c and nodes/list.c compiled where they lie into oracle/_ref/libmotion_ref.so (oracle/ref_motion.c is the
driver), one OS process per QE.

What is held here: (1) the layer registers through the reference's RegisterIPCLayerImpl next to stand-ins for tcp / udpifc / proxy
and is selected by name - no core patch; (2) tuples the reference serialises on the senders come out of the reference's
deserialiser on the receivers bit for bit - fixed-width, NULLs, short and 4-byte-header varlenas, tuples larger than a chunk
(TC_PARTIAL_* chunks across packets) - under hash routing, broadcast and many-to-one; (3) end of stream from every sender ends
the receive loop; (4) a receiver that stops early (LIMIT: SendStopMessage) stops its senders without deadlock."""
import ctypes as C
import multiprocessing as mp
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("CB_MOTION_REF_LIB") or os.path.join(ROOT, "oracle", "_ref", "libmotion_ref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libmotion_ref.so is built from /root/reference (make -C oracle ref)")

INT4, INT8, TEXT = 23, 20, 25
#        typid typlen byval align storage
ATTS = [(INT4, 4, 1, ord("i"), ord("p")), (INT8, 8, 1, ord("d"), ord("p")), (TEXT, -1, 0, ord("i"), ord("x")), (INT4, 4, 1, ord("i"), ord("p"))]


def lib():
    L = C.CDLL(LIB)
    L.ref_motion_run.restype = C.c_int64
    L.ref_aocs_last_error.restype = C.c_char_p
    return L


def rows_of(sender, n, big_every):
    """deterministic rows of one sender: (key, payload, text, nullable int)"""
    rng = np.random.RandomState(1000 + sender)
    out = []
    for i in range(n):
        key = int(rng.randint(0, 1 << 30))
        ln = int(rng.choice([0, 1, 5, 100, 126, 127, 128, 300])) if (big_every == 0 or i % big_every) else 20000   # > one chunk
        text = bytes(rng.randint(32, 127, size=ln, dtype=np.uint8))
        out.append((key, sender * (1 << 40) + i, text, None if i % 7 == 3 else i))
    return out


def encode(rows):
    natts = len(ATTS)
    vals = np.zeros((len(rows), natts), dtype=np.int64)
    nulls = np.zeros((len(rows), natts), dtype=np.uint8)
    var = bytearray()
    for r, row in enumerate(rows):
        for a, v in enumerate(row):
            if v is None:
                nulls[r, a] = 1
            elif ATTS[a][1] == -1:
                vals[r, a] = len(var)
                var += (((len(v) + 4) << 2)).to_bytes(4, "little") + v          # 4-byte varlena header (SET_VARSIZE)
                var += b"\0" * (-len(var) % 4)
            else:
                vals[r, a] = v
    return vals, nulls, bytes(var) or b"\0"


def decode(vals, nulls, var, n):
    out = []
    for r in range(n):
        row = []
        for a in range(len(ATTS)):
            if nulls[r, a]:
                row.append(None)
            elif ATTS[a][1] == -1:
                o = int(vals[r, a])
                if var[o] & 1:                      # 1-byte header (VARATT_IS_1B): length includes the header
                    ln = var[o] >> 1
                    row.append(bytes(var[o + 1:o + ln]))
                else:
                    ln = int.from_bytes(var[o:o + 4], "little") >> 2
                    row.append(bytes(var[o + 4:o + ln]))
            else:
                v = int(vals[r, a])
                row.append(v if ATTS[a][1] == 8 else int(np.int32(v & 0xffffffff)))
        out.append(tuple(row))
    return out


def process(role, idx, nsend, nrecv, session, command, nrows, mode, big_every, stop_after, conn):
    try:
        L = lib()
        natts = len(ATTS)
        arr = lambda k: (C.c_int * natts)(*[a[k] for a in ATTS])      # noqa: E731
        if role == 0:
            rows = rows_of(idx, nrows, big_every)
            vals, nulls, var = encode(rows)
            if mode == "hash":
                routes = np.array([r[0] % nrecv for r in rows], dtype=np.int16)
            elif mode == "broadcast":
                routes = np.full(len(rows), L.ref_motion_broadcast_route(), dtype=np.int16)
            else:
                routes = np.zeros(len(rows), dtype=np.int16)
            n = L.ref_motion_run(0, idx, nsend, nrecv, session, command, natts, arr(0), arr(1), arr(2), arr(3), arr(4),
                                 vals.ctypes.data_as(C.c_void_p), var, nulls.ctypes.data_as(C.c_void_p), C.c_int64(len(rows)),
                                 routes.ctypes.data_as(C.c_void_p), None, None, None, C.c_int64(0), None, C.c_int64(0), C.c_int64(-1))
            conn.send(("sent", idx, int(n), L.ref_aocs_last_error().decode() if n == -1 else ""))
        else:
            cap = nsend * nrows + 16
            vals = np.zeros((cap, natts), dtype=np.int64)
            nulls = np.zeros((cap, natts), dtype=np.uint8)
            src = np.zeros(cap, dtype=np.int16)
            var = C.create_string_buffer(64 << 20)
            n = L.ref_motion_run(1, idx, nsend, nrecv, session, command, natts, arr(0), arr(1), arr(2), arr(3), arr(4), None, None, None,
                                 C.c_int64(0), None, vals.ctypes.data_as(C.c_void_p), nulls.ctypes.data_as(C.c_void_p),
                                 src.ctypes.data_as(C.c_void_p), C.c_int64(cap), var, C.c_int64(64 << 20), C.c_int64(stop_after))
            got = decode(vals, nulls, var.raw, int(n)) if n >= 0 else []
            conn.send(("recv", idx, int(n), got if n >= 0 else L.ref_aocs_last_error().decode()))
    except BaseException as e:      # noqa: BLE001
        conn.send(("error", idx, -99, repr(e)))


_cmd = [os.getpid() % 100000]


def run(nsend, nrecv, nrows, mode, big_every=0, stop_after=-1):
    _cmd[0] += 1
    ctx = mp.get_context("spawn")
    procs, pipes = [], []
    for role, n in ((1, nrecv), (0, nsend)):
        for i in range(n):
            a, b = ctx.Pipe(duplex=False)
            p = ctx.Process(target=process, args=(role, i, nsend, nrecv, os.getpid() % 30000 + 1, _cmd[0], nrows, mode, big_every, stop_after, b), daemon=True)
            p.start()
            procs.append(p)
            pipes.append(a)
    res = [a.recv() if a.poll(120) else ("error", -1, -98, "timed out") for a in pipes]
    for p in procs:
        p.join(10)
        if p.is_alive():
            p.kill()
    return res


def test_layer_registers_through_the_reference_registry():
    L = lib()
    assert L.ref_motion_register() == 0
    # the registry is now full (4 of MAX_NUMBER_TYPES): the reference refuses a fifth
    assert L.ref_motion_register_refusals() == 1


@pytest.mark.parametrize("nsend,nrecv,mode", [(1, 1, "hash"), (2, 2, "hash"), (3, 2, "hash"), (2, 3, "broadcast"), (3, 1, "gather")])
def test_tuples_round_trip_through_the_layer(nsend, nrecv, mode):
    nrows = 1500
    res = run(nsend, nrecv, nrows, mode, big_every=97)
    assert all(r[0] != "error" and r[2] >= 0 for r in res), [r for r in res if r[0] == "error" or r[2] < 0]
    sent = {r[1]: r[2] for r in res if r[0] == "sent"}
    assert sent == {s: nrows for s in range(nsend)}
    want = {d: [] for d in range(nrecv)}
    for s in range(nsend):
        for row in rows_of(s, nrows, 97):
            for d in (range(nrecv) if mode == "broadcast" else [row[0] % nrecv if mode == "hash" else 0]):
                want[d].append(row)
    for r in res:
        if r[0] == "recv":
            assert sorted(r[3], key=lambda t: t[1]) == sorted(want[r[1]], key=lambda t: t[1]), "receiver %d" % r[1]
            # rows of one sender arrive in the order it sent them (a connection is ordered, as TCP / UDPIFC's are)
            for s in range(nsend):
                seq = [t[1] for t in r[3] if t[1] >> 40 == s]
                assert seq == sorted(seq)


def test_receiver_stops_early():
    """LIMIT above the Motion: after 50 rows the receiver sends stop messages; the senders see STOP_SENDING (or finish) and
    everybody terminates"""
    res = run(2, 1, 200000, "gather", stop_after=50)
    assert all(r[0] != "error" for r in res), res
    rc = [r for r in res if r[0] == "recv"][0]
    assert rc[2] == 50
    for r in res:
        if r[0] == "sent":
            assert r[2] <= -10 or r[2] == 200000, r         # STOP_SENDING at some row, or it had already sent everything
    assert any(r[0] == "sent" and r[2] <= -10 for r in res)
