"""The serial half of the device Zstandard decoder (cloudberry_b200/csrc/zstd_dec.cuh: frame / block / literals headers,
FSE and Huffman table construction, Huffman streams, sequence decoding with repeat offsets, XXH64) compiled for the host
and checked against a real libzstd -- the one bundled with pyarrow; the reference links the system's
(gpcontrib/zstd/zstd_compression.c:104-175).  The warp-parallel half is covered on the device by tests/test_gpu_aocs.py;
its batch schedule is replayed here."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

pa = pytest.importorskip("pyarrow")

from oracle import aocs_format as A  # noqa: E402
from test_aocs_format import ZSTDCASES  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
STATS = ("blk_raw blk_rle blk_cmp lit_raw lit_rle lit_cmp lit_treeless s4 s1 w_direct w_fse LL0 LL1 LL2 LL3 OF0 OF1 OF2 OF3 "
         "ML0 ML1 ML2 ML3 multiblock nseq0 longnseq").split()


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("zstd") / "libzstdhost.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-o", so, os.path.join(HERE, "native", "zstd_host.cpp")])
    L = C.CDLL(so)
    L.zstd_host_decompress.restype = C.c_longlong
    L.zstd_host_decompress.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint32]
    L.zstd_host_xxh64.restype = C.c_ulonglong
    L.zstd_host_xxh64.argtypes = [C.c_char_p, C.c_ulonglong]
    L.zstd_host_stat.restype = C.c_ulonglong
    return L


def unzstd(L, z, cap):
    out = (C.c_ubyte * max(cap, 1))()
    r = L.zstd_host_decompress(z, len(z), out, cap)
    return r, bytes(out[:max(r, 0)])


def corpus():
    rng = np.random.default_rng(5)
    for t in range(330):
        n = int(rng.integers(1, 400000 if t % 10 == 0 else 60000))
        kind = t % 11
        if kind == 0:
            src = rng.integers(0, 256, n, dtype=np.uint8)
        elif kind == 1:
            src = rng.integers(0, 4, n, dtype=np.uint8)
        elif kind == 2:
            src = ((np.arange(n) // 7) % 251).astype(np.uint8)
        elif kind == 3:
            src = np.where(np.arange(n) % 8 < 2, rng.integers(0, 256, n), 0).astype(np.uint8)
        elif kind == 4:
            src = np.frombuffer(b"abcdefgh ijk", dtype=np.uint8)[rng.integers(0, 12, n)]
        elif kind == 5:
            src = np.repeat(rng.integers(0, 256, n // 50 + 1, dtype=np.uint8), 50)[:n]
        elif kind == 6:
            src = np.cumsum(rng.integers(0, 3, n // 4 + 1)).astype(np.int32).view(np.uint8)[:n]
        elif kind == 7:
            src = np.full(n, int(rng.integers(0, 256)), dtype=np.uint8)                       # RLE blocks
        elif kind == 8:
            # the same literal byte between matches of varying length -> RLE literals
            src = np.concatenate([np.concatenate([np.zeros(int(k), np.uint8), np.arange(int(m), dtype=np.uint8) % 3 + 1])
                                  for k, m in zip(rng.integers(1, 40, 300), rng.integers(4, 30, 300))])
        elif kind == 9:
            src = rng.normal(0, 30, n // 8 + 1).round().astype(np.float64).view(np.uint8)[:n]
        else:
            src = np.tile(rng.integers(0, 256, 5000, dtype=np.uint8), n // 5000 + 1)[:n]       # far matches
        src = src.tobytes()
        level = (1, 3, 5, 9, 15, 19, -3)[t % 7]
        yield src, pa.Codec("zstd", compression_level=level).compress(src, asbytes=True)
    # several blocks with similar statistics -> Repeat mode tables, treeless literals
    n = 400000
    for src, level in ((np.repeat(rng.integers(0, 256, n // 50 + 1, dtype=np.uint8), 50)[:n], 19),
                       (np.where(np.arange(n) % 8 < 2, rng.integers(0, 256, n), 0).astype(np.uint8), 9),
                       (np.frombuffer(b"abcdefgh ijk", dtype=np.uint8)[rng.integers(0, 12, n)], 15),
                       (rng.integers(0, 4, n, dtype=np.uint8), 3)):
        src = src.tobytes()
        yield src, pa.Codec("zstd", compression_level=level).compress(src, asbytes=True)
    # a block whose literals are all the same byte (pieces of an earlier random area glued by 'x') -> RLE literals
    d = rng.integers(0, 256, 100000, dtype=np.uint8)
    d[d == ord("x")] = ord("y")
    parts = [d, d[:31072]]
    for o, m in zip(rng.integers(0, 90000, 6000), rng.integers(20, 60, 6000)):
        parts.append(d[int(o):int(o) + int(m)])
        parts.append(np.frombuffer(b"x", dtype=np.uint8))
    src = np.concatenate(parts).tobytes()
    for level in (1, 9, 19):
        yield src, pa.Codec("zstd", compression_level=level).compress(src, asbytes=True)


def test_xxh64_known_answers(lib):
    assert lib.zstd_host_xxh64(b"", 0) == 0xEF46DB3751D8E999
    assert lib.zstd_host_xxh64(b"abc", 3) == 0x44BC2CF5AD770999
    assert lib.zstd_host_xxh64(b"Nobody inspects the spammish repetition", 39) == 0xFBCEA83C8A378BF1


def test_against_libzstd(lib):
    for src, z in corpus():
        r, out = unzstd(lib, z, len(src))
        assert r == len(src) and out == src
    seen = {n: lib.zstd_host_stat(i) for i, n in enumerate(STATS)}
    # the corpus reaches every block type, literals type, weight encoding and table mode of the format
    missing = [n for n in STATS if seen[n] == 0]
    assert not missing, (missing, seen)


def test_batch_schedule_of_the_warp_kernel(lib):
    lib.zstd_host_set_warp_mode(1)
    try:
        for src, z in corpus():
            r, out = unzstd(lib, z, len(src))
            assert r == len(src) and out == src
    finally:
        lib.zstd_host_set_warp_mode(0)


def test_frame_checksum(lib):
    """a frame with Content_Checksum_flag: pyarrow does not write one, so set the flag and append the XXH64 low word by
    hand (the hash itself is pinned by the known answers above)"""
    src = bytes(range(256)) * 300
    z = bytearray(pa.Codec("zstd", compression_level=3).compress(src, asbytes=True))
    z[4] |= 0x04
    good = bytes(z) + (lib.zstd_host_xxh64(src, len(src)) & 0xFFFFFFFF).to_bytes(4, "little")
    assert unzstd(lib, good, len(src)) == (len(src), src)
    assert unzstd(lib, good[:-1] + bytes([good[-1] ^ 1]), len(src))[0] < 0
    assert unzstd(lib, bytes(z), len(src))[0] < 0                                              # checksum missing


def test_bad_frames_are_refused_not_followed(lib):
    src = (b"the quick brown fox " * 3000)[:50000] + bytes(np.random.default_rng(2).integers(0, 256, 20000, dtype=np.uint8))
    z = pa.Codec("zstd", compression_level=5).compress(src, asbytes=True)
    assert unzstd(lib, z, len(src))[0] == len(src)
    assert unzstd(lib, z, len(src) - 1)[0] < 0                       # longer than the block header's dataLength
    assert unzstd(lib, z[:len(z) // 2], len(src))[0] < 0             # truncated
    assert unzstd(lib, b"\x27" + z[1:], len(src))[0] < 0            # magic
    assert unzstd(lib, z[:4] + bytes([z[4] | 0x08]) + z[5:], len(src))[0] < 0     # reserved bit
    assert unzstd(lib, z[:4] + bytes([z[4] | 0x01]) + z[5:], len(src))[0] < 0     # dictionary id
    rng = np.random.default_rng(11)
    for t in range(400):
        zb = bytearray(z)
        zb[int(rng.integers(4, len(z)))] ^= 1 << int(rng.integers(0, 8))
        r, out = unzstd(lib, bytes(zb), len(src))                      # must come back; without a checksum a flip in
        assert r < 0 or r == len(src)                                  # raw literals goes unnoticed, as with libzstd


@pytest.mark.parametrize("case", ZSTDCASES, ids=[c[0] for c in ZSTDCASES])
def test_reference_written_compressed_blocks(lib, case):
    name, typname, checksum, blocksize, dscale, nblocks, raw, values, nulls, zlevel = case
    for b in A.walk_blocks_ex(raw, checksum):
        if not b["clen"]:
            continue
        z = raw[b["off"]:b["off"] + b["clen"]]
        r, out = unzstd(lib, z, b["dlen"])
        assert r == b["dlen"] and out == pa.Codec("zstd").decompress(z, decompressed_size=b["dlen"], asbytes=True)
