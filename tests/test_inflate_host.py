"""The serial half of the device inflater (cloudberry_b200/csrc/inflate.cuh: block headers, Huffman tables, symbol decoding
into the literal / match queue) compiled for the host and checked against the system zlib, the library the reference
calls (compress2 / uncompress, catalog/pg_compression.c:250-251).  The warp-parallel half (applying the queue, Adler-32)
is covered on the device by tests/test_gpu_aocs.py."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest

from oracle import aocs_format as A
from test_aocs_format import ZCASES

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("infl") / "libinflhost.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-o", so, os.path.join(HERE, "native", "inflate_host.cpp")])
    L = C.CDLL(so)
    L.infl_host_zlib.restype = C.c_longlong
    L.infl_host_zlib.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    L.infl_host_zlib_warp.restype = C.c_longlong
    L.infl_host_zlib_warp.argtypes = L.infl_host_zlib.argtypes
    return L


def inflate(L, z, cap, warp=False):
    out = (C.c_ubyte * max(cap, 1))()
    ad = C.c_uint32()
    r = (L.infl_host_zlib_warp if warp else L.infl_host_zlib)(z, len(z), out, cap, C.byref(ad))
    return r, bytes(out[:max(r, 0)]), ad.value


def streams():
    rng = np.random.default_rng(7)
    for t in range(240):
        n = int(rng.integers(1, 300000 if t % 12 == 0 else 50000))
        kind = t % 6
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
        else:
            src = np.repeat(rng.integers(0, 256, n // 50 + 1, dtype=np.uint8), 50)[:n]
        src = src.tobytes()
        level = (0, 1, 6, 9)[t % 4]
        strategy = {3: zlib.Z_FIXED, 5: zlib.Z_HUFFMAN_ONLY, 6: zlib.Z_RLE}.get(t % 7, zlib.Z_DEFAULT_STRATEGY)
        co = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strategy)
        yield src, co.compress(src) + co.flush()


def test_against_system_zlib(lib):
    for src, z in streams():
        r, out, ad = inflate(lib, z, len(src))
        assert r == len(src) and out == src
        assert ad == zlib.adler32(src)


def test_batch_schedule_of_the_warp_kernel(lib):
    """the order in which k_aocs_inflate's warp applies a batch of 32 entries (independent entries in any order, then the
    self-referencing ones in sequence, 32 bytes at a time) gives the serial result"""
    for src, z in streams():
        r, out, ad = inflate(lib, z, len(src), warp=True)
        assert r == len(src) and out == src


def test_sync_flushes_and_many_blocks(lib):
    """streams with empty stored blocks (Z_SYNC_FLUSH), full flushes, and mixed block types back to back"""
    rng = np.random.default_rng(9)
    co = zlib.compressobj(6)
    parts, src = [], b""
    for i in range(40):
        chunk = rng.integers(0, 8 if i % 3 else 256, int(rng.integers(0, 5000)), dtype=np.uint8).tobytes()
        src += chunk
        parts.append(co.compress(chunk))
        parts.append(co.flush(zlib.Z_SYNC_FLUSH if i % 2 else zlib.Z_FULL_FLUSH))
    z = b"".join(parts) + co.flush()
    r, out, ad = inflate(lib, z, len(src))
    assert r == len(src) and out == src and ad == zlib.adler32(src)


def test_bad_streams_are_refused_not_followed(lib):
    src = bytes(range(256)) * 40
    z = zlib.compress(src, 6)
    assert inflate(lib, z, len(src))[0] == len(src)
    assert inflate(lib, z, len(src) - 1)[0] < 0                     # output larger than the header's dataLength
    assert inflate(lib, z[:len(z) // 2], len(src))[0] < 0           # truncated
    assert inflate(lib, b"\x78\x9d" + z[2:], len(src))[0] < 0       # header check bits
    assert inflate(lib, b"\x78\xbb" + z[2:], len(src))[0] < 0       # preset dictionary flag
    assert inflate(lib, z[:2] + b"\x07" + z[3:], len(src))[0] < 0   # block type 3
    rng = np.random.default_rng(11)
    bad = 0
    for t in range(300):
        zb = bytearray(z)
        zb[int(rng.integers(2, len(z) - 4))] ^= 1 << int(rng.integers(0, 8))
        r, out, ad = inflate(lib, bytes(zb), len(src))
        # a flipped bit is either caught by the decoder, or changes the output so that length or Adler-32 disagree
        if r < 0 or r != len(src) or ad != zlib.adler32(out):
            bad += 1
        else:
            assert out == src                                        # (a flip in unused padding bits)
    assert bad > 250


@pytest.mark.parametrize("case", ZCASES, ids=[c[0] for c in ZCASES])
def test_reference_written_compressed_blocks(lib, case):
    """every bulk-compressed block of the zlib fixtures (written through the reference's header makers)"""
    name, typname, checksum, blocksize, dscale, nblocks, raw, values, nulls, zlevel = case
    for b in A.walk_blocks_ex(raw, checksum):
        if not b["clen"]:
            continue
        z = raw[b["off"]:b["off"] + b["clen"]]
        r, out, ad = inflate(lib, z, b["dlen"])
        assert r == b["dlen"] and out == zlib.decompress(z) and ad == zlib.adler32(out)
