"""Pin the oracle's hash restatement (not gpu): known answers computed from the reference's own
hashfn.o and, where oracle/_ref was built from /root/reference/src/common/hashfn.c,
the compiled reference itself on random inputs."""
import ctypes as C
import random
import struct

import pytest

from cloudberry_b200 import plan as P

KNOWN_U32 = {0: 0xefbec0af, 1: 0x8e731746, 2: 0x439edcf6, 42: 0x59fcfec8, 12345: 0xfb58525d, 0xffffffff: 0x16fe094a}
KNOWN_CHAR = {"A": 0x62bf9fee, "N": 0x65bad44b, "R": 0xf1c13a4e, "F": 0x6fb5dc39, "O": 0xb09a58de}
KNOWN_SEG8 = {0: 1, 1: 4, 2: 3, 42: 0, 12345: 6, 0xffffffff: 2}
KNOWN_SEG3 = {0: 1, 1: 1, 2: 0, 42: 0, 12345: 2, 0xffffffff: 2}


def _s32(k):
    return k - (1 << 32) if k >= (1 << 31) else k


def test_known_answers(oracle):
    L = oracle.lib()
    for k, h in KNOWN_U32.items():
        assert L.ora_hash_datum(P.INT4, _s32(k)) == h
        assert L.ora_hash_datum(P.DATE, _s32(k)) == h          # date hashes as int4 (pg_amproc.dat:310)
    for ch, h in KNOWN_CHAR.items():
        assert L.ora_hash_datum(P.BPCHAR1, ord(ch)) == h
        assert oracle.hashbpchar(ch) == h
        assert oracle.hashbpchar(ch + "    ") == h             # bcTruelen strips blanks (varchar.c:997)
    bits = struct.unpack("<q", struct.pack("<d", 1.5))[0]
    assert L.ora_hash_datum(P.FLOAT8, bits) == 0x259a2972
    assert L.ora_hash_datum(P.FLOAT8, struct.unpack("<q", struct.pack("<d", -0.0))[0]) == 0


def test_hashint8_compatible_with_int4(oracle):
    L = oracle.lib()
    for k in [0, 1, -1, 42, -42, 2 ** 31 - 1, -2 ** 31]:
        assert L.ora_hash_datum(P.INT8, k) == L.ora_hash_datum(P.INT4, k)


def test_cdbhash_known_segments(oracle):
    L = oracle.lib()
    t = (C.c_int32 * 1)(P.INT4)
    for k, seg in KNOWN_SEG8.items():
        assert L.ora_cdbhash_segment(t, (C.c_int64 * 1)(_s32(k)), None, 1, 8) == seg
    for k, seg in KNOWN_SEG3.items():
        assert L.ora_cdbhash_segment(t, (C.c_int64 * 1)(_s32(k)), None, 1, 3) == seg


def test_against_compiled_reference(oracle):
    R = oracle.ref_hash_lib()
    if R is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    L = oracle.lib()
    rng = random.Random(7)
    for _ in range(2000):
        k = rng.getrandbits(32)
        assert L.ora_hash_datum(P.INT4, _s32(k)) == R.hash_bytes_uint32(k)
    for n in list(range(0, 40)) + [100, 255]:
        s = bytes(rng.getrandbits(8) | 1 for _ in range(n)).replace(b" ", b"x")
        assert L.ora_hashbpchar_text(s, len(s)) == R.hash_bytes(s, len(s))
    # int8: fold then hash_uint32 (hashfunc.c:84-102)
    for _ in range(500):
        v = rng.getrandbits(64) - (1 << 63)
        lo = v & 0xffffffff
        hi = (v >> 32) & 0xffffffff
        lo ^= hi if v >= 0 else (~hi & 0xffffffff)
        assert L.ora_hash_datum(P.INT8, v) == R.hash_bytes_uint32(lo)
    # float8: hash_any over the 8 bytes
    for _ in range(200):
        f = rng.uniform(-1e9, 1e9)
        bits = struct.unpack("<q", struct.pack("<d", f))[0]
        assert L.ora_hash_datum(P.FLOAT8, bits) == R.hash_bytes(struct.pack("<d", f), 8)


def test_numeric_avg_text_rule(oracle):
    """select_div_scale worked examples from SURVEY.md (golden Q1 row A/F)."""
    L = oracle.lib()
    buf = C.create_string_buffer(128)
    for s, ds, n, want in [(38045600, 2, 14876, "25.5751546114546921"), (53234821165, 2, 14876, "35785.709306937349"),
                           (74501, 2, 14876, "0.05008133906964237698"), (0, 2, 5, "0.00000000000000000000"),
                           (-38045600, 2, 14876, "-25.5751546114546921"), (10, 0, 4, "2.5000000000000000"),
                           (1, 0, 3, "0.33333333333333333333"), (2, 0, 3, "0.66666666666666666667")]:
        L.ora_numeric_avg_text(s & (2 ** 64 - 1) if s >= 0 else s, -1 if s < 0 else 0, ds, n, buf, 128)
        assert buf.value.decode() == want, (s, ds, n)
