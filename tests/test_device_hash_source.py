"""The hash functions the kernels inline, as SOURCE (cloudberry_b200/csrc/common.cuh, __host__ __device__), compiled for the host
and held against the reference's own code (not gpu): hashint4 / hashint8 / hashfloat8 / hashbpchar, murmurhash32, the
rotate-xor key combination of nodeHash.c / execGrouping.c / cdbhash.c, and jump_consistent_hash through cdbhashreduce.
On the device the same text runs with __ddiv_rn / __dmul_rn in the jump hash (IEEE, no contraction), which
tests/test_gpu_multirank.py and the Motion tests check against the oracle's routes."""
import ctypes as C
import os
import random
import shutil
import struct
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "..", "oracle", "_ref", "libexec_ref.so")
INT4, INT8, FLOAT8, BPCHAR, TEXT = range(5)


@pytest.fixture(scope="module")
def libs(tmp_path_factory):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libexec_ref.so not built (no /root/reference on this box)")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    so = str(tmp_path_factory.mktemp("hh") / "libhashhost.so")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "--fmad=false", "-shared",
                           "-Xcompiler", "-fPIC", "-o", so, os.path.join(HERE, "native", "hash_host.cu")])
    H = C.CDLL(so)
    for name, args in (("hh_hash_uint32", [C.c_uint32]), ("hh_hashint8", [C.c_int64]), ("hh_hashfloat8", [C.c_uint64]),
                       ("hh_hash_bpchar1", [C.c_uint8]), ("hh_hash_bytes", [C.c_char_p, C.c_int]), ("hh_murmurhash32", [C.c_uint32]),
                       ("hh_hash_combine", [C.c_uint32, C.c_uint32, C.c_int])):
        getattr(H, name).restype = C.c_uint32
        getattr(H, name).argtypes = args
    H.hh_jump_consistent_hash.restype = C.c_int32
    H.hh_jump_consistent_hash.argtypes = [C.c_uint64, C.c_int32]
    R = C.CDLL(REF)
    R.ref_hash_datum.restype = C.c_uint32
    R.ref_hash_datum.argtypes = [C.c_int, C.c_int64, C.c_char_p, C.c_int]
    R.ref_cdbhash_segment.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_char_p),
                                      C.POINTER(C.c_uint8), C.c_int]
    R.ref_jump_consistent_hash.argtypes = [C.c_uint32, C.c_int]
    P = C.CDLL(os.path.join(HERE, "..", "oracle", "_ref", "libpg_hashfn.so"))
    P.ref_murmurhash32.restype = C.c_uint32
    P.ref_murmurhash32.argtypes = [C.c_uint32]
    return H, R, P


def _bits(f):
    return struct.unpack("<Q", struct.pack("<d", f))[0]


def _s64(u):
    return u - (1 << 64) if u >= (1 << 63) else u


def test_type_hash_functions(libs):
    H, R, _ = libs
    rng = random.Random(77)
    for k in [0, 1, 42, 0x7fffffff, 0x80000000, 0xffffffff] + [rng.getrandbits(32) for _ in range(5000)]:
        s = k - (1 << 32) if k >= (1 << 31) else k
        assert H.hh_hash_uint32(k) == R.ref_hash_datum(INT4, s, None, 0)
    for v in [0, 1, -1, 2 ** 31, -2 ** 31 - 1, 2 ** 63 - 1, -2 ** 63] + [rng.getrandbits(64) - 2 ** 63 for _ in range(5000)]:
        assert H.hh_hashint8(v) == R.ref_hash_datum(INT8, v, None, 0), v
    fl = [0.0, -0.0, 1.5, float("inf"), float("-inf"), float("nan"), 5e-324, 1.7e308] + [rng.uniform(-1e15, 1e15) for _ in range(3000)]
    for f in fl:
        assert H.hh_hashfloat8(_bits(f)) == R.ref_hash_datum(FLOAT8, _s64(_bits(f)), None, 0), f
    for nan_bits in (0xfff8000000000001, 0x7ff0000000000001, 0xffffffffffffffff):
        assert H.hh_hashfloat8(nan_bits) == R.ref_hash_datum(FLOAT8, _s64(nan_bits), None, 0)
    for ch in range(1, 256):
        assert H.hh_hash_bpchar1(ch) == R.ref_hash_datum(BPCHAR, 0, bytes([ch]), 1), ch
    for n in list(range(0, 40)) + [63, 64, 65, 500]:
        s = bytes(rng.randrange(33, 127) for _ in range(n))
        assert H.hh_hash_bytes(s, n) == R.ref_hash_datum(TEXT, 0, s, n)


def test_murmur_and_key_combination(libs):
    H, R, P = libs
    rng = random.Random(78)
    for _ in range(5000):
        x = rng.getrandbits(32)
        assert H.hh_murmurhash32(x) == P.ref_murmurhash32(x)
    # the Motion's use of the combination: cdbhash over several keys, NULLs included, then the jump hash
    kinds = [INT4, INT8, FLOAT8]
    for _ in range(4000):
        nk = rng.randrange(1, 5)
        nseg = rng.choice([1, 2, 3, 4, 8, 16, 48, 64, 1000])
        ks = [rng.choice(kinds) for _ in range(nk)]
        nulls = [1 if rng.random() < 0.15 else 0 for _ in range(nk)]
        vals = []
        acc = 0
        for k, nul in zip(ks, nulls):
            if k == INT4:
                v = rng.getrandbits(32) - 2 ** 31
                h = H.hh_hash_uint32(v & 0xffffffff)
            elif k == INT8:
                v = rng.getrandbits(64) - 2 ** 63
                h = H.hh_hashint8(v)
            else:
                b = _bits(rng.uniform(-1e9, 1e9))
                v = _s64(b)
                h = H.hh_hashfloat8(b)
            vals.append(v)
            acc = H.hh_hash_combine(acc, h, nul)
        want = R.ref_cdbhash_segment(nk, (C.c_int * nk)(*ks), (C.c_int64 * nk)(*vals), None, (C.c_uint8 * nk)(*nulls), nseg)
        assert want >= 0
        assert H.hh_jump_consistent_hash(acc, nseg) == want, (ks, vals, nulls, nseg)
        assert R.ref_jump_consistent_hash(acc, nseg) == want
