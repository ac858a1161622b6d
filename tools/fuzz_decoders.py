#!/usr/bin/env python3
"""AddressSanitizer + UBSan fuzzing of the host builds of the decoders that also run on the device: the serial zlib and
Zstandard decoders (cloudberry_b200/csrc/inflate.cuh, zstd_dec.cuh through tests/native/*.cpp, exact-size heap buffers so
any read past the input or write past the output is reported), the tuple chunk parser (csrc/exec/cb_tupser.c) and the numeric
finaliser (csrc/exec/cb_numeric.c: 128-bit states at their limits, every display scale, output buffers that are too small).
Streams are bit-flipped, truncated and given too-small outputs.  CPU only:

    python tools/fuzz_decoders.py [rounds]

Re-executes itself with libasan preloaded.  DESIGN.md 8 quotes the last run."""
import ctypes as C
import os
import subprocess
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = "/tmp/cbgpu_fuzz"


def build():
    os.makedirs(OUT, exist_ok=True)
    flags = ["-g", "-O1", "-fsanitize=address,undefined", "-shared", "-fPIC"]
    subprocess.check_call(["g++", "-std=c++17"] + flags + ["-o", OUT + "/libz.so", os.path.join(ROOT, "tests", "native", "zstd_host.cpp")])
    subprocess.check_call(["g++", "-std=c++17"] + flags + ["-o", OUT + "/libi.so", os.path.join(ROOT, "tests", "native", "inflate_host.cpp")])
    subprocess.check_call(["gcc"] + flags + ["-I" + os.path.join(ROOT, "include"), "-o", OUT + "/libt.so",
                                             os.path.join(ROOT, "cloudberry_b200", "csrc", "exec", "cb_tupser.c")])
    subprocess.check_call(["gcc"] + flags + ["-fno-sanitize-recover=undefined", "-I" + os.path.join(ROOT, "include"), "-o", OUT + "/libn.so",
                                             os.path.join(ROOT, "cloudberry_b200", "csrc", "exec", "cb_numeric.c")])


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    if "libasan" not in os.environ.get("LD_PRELOAD", ""):
        build()
        asan = subprocess.check_output(["gcc", "-print-file-name=libasan.so"], text=True).strip()
        env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0")
        sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__), str(rounds)], env=env))
    import numpy as np
    import pyarrow as pa
    Z, Inf, T = C.CDLL(OUT + "/libz.so"), C.CDLL(OUT + "/libi.so"), C.CDLL(OUT + "/libt.so")
    Z.zstd_host_decompress.restype = C.c_longlong
    Z.zstd_host_decompress.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    for f in (Inf.infl_host_zlib, Inf.infl_host_zlib_warp):
        f.restype = C.c_longlong
        f.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    libc.free.argtypes = [C.c_void_p]

    def run(fn, z, cap, extra=()):
        pin = libc.malloc(max(len(z), 1))
        C.memmove(pin, z, len(z))
        pout = libc.malloc(max(cap, 1))
        r = fn(pin, len(z), pout, cap, *extra)
        libc.free(pin)
        libc.free(pout)
        return r
    rng = np.random.default_rng(1)
    n = 0
    ad = C.c_uint32()
    for t in range(rounds):
        m = int(rng.integers(1, 40000))
        k = t % 5
        src = (rng.integers(0, 256, m, dtype=np.uint8) if k == 0 else rng.integers(0, 4, m, dtype=np.uint8) if k == 1
               else ((np.arange(m) // 7) % 251).astype(np.uint8) if k == 2 else np.repeat(rng.integers(0, 256, m // 50 + 1, dtype=np.uint8), 50)[:m]
               if k == 3 else np.cumsum(rng.integers(0, 3, m // 4 + 1)).astype(np.int32).view(np.uint8)[:m]).tobytes()
        zs = pa.Codec("zstd", compression_level=[1, 3, 9, 19][t % 4]).compress(src, asbytes=True)
        zl = zlib.compress(src, [1, 6, 9][t % 3])
        assert run(Z.zstd_host_decompress, zs, len(src)) == len(src)
        assert run(Inf.infl_host_zlib, zl, len(src), (C.byref(ad),)) == len(src)
        for j in range(25):
            for z, fn, extra in ((zs, Z.zstd_host_decompress, ()), (zl, Inf.infl_host_zlib, (C.byref(ad),)), (zl, Inf.infl_host_zlib_warp, (C.byref(ad),))):
                zb = bytearray(z)
                for _ in range(int(rng.integers(1, 4))):
                    zb[int(rng.integers(0, len(zb)))] ^= 1 << int(rng.integers(0, 8))
                cut = len(zb) if j % 5 else int(rng.integers(0, len(zb) + 1))
                cap = len(src) if j % 7 else int(rng.integers(0, len(src) + 1))
                run(fn, bytes(zb[:cut]), cap, extra)
                n += 1
    print("decoders: %d damaged streams, no sanitizer report" % n)
    # tuple chunk parser: serialise random rows, damage the stream, parse from an exact-size buffer
    class Attr(C.Structure):
        _fields_ = [("type", C.c_int32), ("dscale", C.c_int32), ("bpchar_len", C.c_int32), ("ntexts", C.c_int32),
                    ("texts", C.POINTER(C.c_char_p)), ("text_lens", C.POINTER(C.c_int32))]
    T.cb_tupser_row.restype = C.c_int64
    T.cb_tupser_row.argtypes = [C.POINTER(Attr), C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_uint8), C.c_int, C.c_void_p, C.c_int64]
    T.cb_tupser_next.restype = C.c_int64
    T.cb_tupser_next.argtypes = [C.POINTER(Attr), C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_uint8)]
    texts = (C.c_char_p * 3)(b"AIR", b"MAIL", b"zzz")
    lens = (C.c_int32 * 3)(3, 4, 3)
    at = (Attr * 6)()
    for i, (ty, ds, bl) in enumerate(((1, 0, 0), (4, 2, 0), (7, 0, 10), (2, 0, 0), (9, 0, 0), (5, 0, 0))):   # cb_plan.h CbTypeId values
        at[i].type, at[i].dscale, at[i].bpchar_len = ty, ds, bl
    at[2].ntexts, at[2].texts, at[2].text_lens = 3, texts, lens
    buf = C.create_string_buffer(8192)
    vals, nl = (C.c_int64 * 6)(), (C.c_uint8 * 6)()
    out_v, out_n, used = (C.c_int64 * 6)(), (C.c_uint8 * 6)(), C.c_int64()
    m = 0
    for t in range(rounds * 50):
        row = [int(rng.integers(-2**31, 2**31)), int(rng.integers(-2**50, 2**50)), int(rng.integers(0, 3)), int(rng.integers(-2**62, 2**62)),
               int(rng.integers(0, 2)), int(rng.integers(65, 91))]
        for i in range(6):
            vals[i] = row[i]
            nl[i] = int(rng.random() < 0.2)
        k = T.cb_tupser_row(at, 6, vals, nl, int(rng.choice([8160, 24, 40, 64])), buf, len(buf))
        assert k > 0
        data = bytearray(buf.raw[:k])
        for _ in range(int(rng.integers(0, 4))):
            data[int(rng.integers(0, len(data)))] ^= 1 << int(rng.integers(0, 8))
        cut = len(data) if t % 4 else int(rng.integers(0, len(data) + 1))
        pin = libc.malloc(max(cut, 1))
        C.memmove(pin, bytes(data[:cut]), cut)
        pos = 0
        for _ in range(8):
            rc = T.cb_tupser_next(at, 6, pin + pos, cut - pos, C.byref(used), out_v, out_n)
            m += 1
            if rc != 1:
                break
            pos += used.value
        libc.free(pin)
    print("tuple chunks: %d parses of damaged streams, no sanitizer report" % m)
    # numeric finalisation: extreme (sum, N) states into exact-size output buffers
    import random
    N = C.CDLL(OUT + "/libn.so")
    N.cb_numeric_sum_text.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int32]
    N.cb_numeric_avg_text.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_int32]
    pr = random.Random(1)
    k = 0
    for _ in range(rounds * 100):
        ds = pr.choice([0, 1, 2, 4, 6, 9, 12, 31])
        v = pr.choice([0, 1, -1, 2 ** 127 - 1, -2 ** 127, pr.randrange(-10 ** 38, 10 ** 38), pr.randrange(-10 ** 6, 10 ** 6)])
        cnt = pr.choice([1, 2, 3, 7, 2 ** 31, 2 ** 62, 2 ** 63 - 1, pr.randrange(1, 10 ** 12)])
        u = v & (2 ** 128 - 1)
        lo, hi = u & (2 ** 64 - 1), u >> 64
        lo, hi = (lo - 2 ** 64 if lo >= 2 ** 63 else lo), (hi - 2 ** 64 if hi >= 2 ** 63 else hi)
        for cap in (200, 48, 8, 1):
            pout = libc.malloc(cap)
            N.cb_numeric_sum_text(lo, hi, ds, pout, cap)
            N.cb_numeric_avg_text(lo, hi, ds, cnt, pout, cap)
            libc.free(pout)
            k += 2
    print("numeric finaliser: %d calls at the limits of the state, no sanitizer report" % k)

if __name__ == "__main__":
    main()
