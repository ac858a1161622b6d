"""The oracle's restatements and the product's host finalisation against the REFERENCE's own code (not gpu).

oracle/_ref/libexec_ref.so = the reference's hashfunc.c, varchar.c, cdbhash.c and numeric.c compiled where they lie
(oracle/Makefile, oracle/ref_exec.c): hash support functions, Motion hashing (makeCdbHash / cdbhash / cdbhashreduce with
jump_consistent_hash) and the numeric transition / final functions of sum and avg, called as nodeAgg.c / nodeMotion.c call
them.  These tests pin, on random inputs,
  * oracle/pg_hash.h (ora_hash_datum, ora_hashbpchar_text) and ora_cdbhash_segment - the values every GPU parity test
    compares the device's hashes, join buckets, group hashes and Motion routes with;
  * cb_numeric_sum_text / cb_numeric_avg_text (csrc/exec/cb_numeric.c, product) and ora_numeric_*_text (oracle): the
    reference's numeric text of sum / avg from the exact (sum, N) pair the device delivers;
  * the scaled-integer rule for l_extendedprice * (1 - l_discount) * (1 + l_tax): numeric_sub / numeric_mul add display
    scales exactly as the int64 pipeline assumes.
Skipped where oracle/_ref was not built (no /root/reference on the box and no prebuilt copy)."""
import ctypes as C
import os
import random
import struct

import pytest

from cloudberry_b200 import capi
from cloudberry_b200 import plan as P

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "..", "oracle", "_ref", "libexec_ref.so")

INT4, INT8, FLOAT8, BPCHAR, TEXT = range(5)


@pytest.fixture(scope="module")
def ref(oracle):
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref/libexec_ref.so not built (no /root/reference on this box)")
    R = C.CDLL(SO)
    R.ref_exec_last_error.restype = C.c_char_p
    R.ref_hash_datum.restype = C.c_uint32
    R.ref_hash_datum.argtypes = [C.c_int, C.c_int64, C.c_char_p, C.c_int]
    R.ref_cdbhash_segment.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_char_p),
                                      C.POINTER(C.c_uint8), C.c_int]
    R.ref_jump_consistent_hash.argtypes = [C.c_uint32, C.c_int]
    R.ref_numeric_binop.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    R.ref_numeric_agg.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_char_p, C.c_int]
    R.ref_int8_agg.argtypes = [C.POINTER(C.c_int64), C.c_int, C.c_char_p, C.c_char_p, C.c_int]
    R.ref_int4_sum.argtypes = [C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int64)]
    return R


def _bits(f):
    return struct.unpack("<q", struct.pack("<d", f))[0]


def test_hash_support_functions(oracle, ref):
    """hashint4 / hashint8 / hashfloat8 / hashbpchar / hashtext as compiled from the reference == oracle/pg_hash.h"""
    L = oracle.lib()
    rng = random.Random(11)
    ints = [0, 1, -1, 42, 2 ** 31 - 1, -2 ** 31] + [rng.getrandbits(32) - 2 ** 31 for _ in range(3000)]
    for k in ints:
        assert L.ora_hash_datum(P.INT4, k) == ref.ref_hash_datum(INT4, k, None, 0)
        assert L.ora_hash_datum(P.DATE, k) == ref.ref_hash_datum(INT4, k, None, 0)
    bigs = [0, 1, -1, 2 ** 31, -2 ** 31 - 1, 2 ** 32, -2 ** 32, 2 ** 63 - 1, -2 ** 63] + \
           [rng.getrandbits(64) - 2 ** 63 for _ in range(3000)]
    for v in bigs:
        assert L.ora_hash_datum(P.INT8, v) == ref.ref_hash_datum(INT8, v, None, 0), v
    floats = [0.0, -0.0, 1.5, -1.5, float("inf"), float("-inf"), float("nan"), 1e-310, 1e308] + \
             [rng.uniform(-1e12, 1e12) for _ in range(2000)]
    for f in floats:
        assert L.ora_hash_datum(P.FLOAT8, _bits(f)) == ref.ref_hash_datum(FLOAT8, _bits(f), None, 0), f
    # a NaN with another payload / sign hashes like the canonical one (hashfunc.c:205-213)
    other_nan = struct.unpack("<q", struct.pack("<Q", 0xfff8000000000123))[0]
    assert L.ora_hash_datum(P.FLOAT8, other_nan) == ref.ref_hash_datum(FLOAT8, other_nan, None, 0)
    assert ref.ref_hash_datum(FLOAT8, other_nan, None, 0) == ref.ref_hash_datum(FLOAT8, _bits(float("nan")), None, 0)
    for n in list(range(0, 48)) + [100, 255, 1000]:
        s = bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789 -#") for _ in range(n))
        pad = s + b" " * rng.randrange(0, 12)
        # character(n): trailing blanks do not count (bcTruelen); text / varchar: they do
        assert L.ora_hashbpchar_text(pad, len(pad)) == ref.ref_hash_datum(BPCHAR, 0, pad, len(pad)), pad
        assert ref.ref_hash_datum(BPCHAR, 0, pad, len(pad)) == ref.ref_hash_datum(BPCHAR, 0, s.rstrip(b" "), len(s.rstrip(b" ")))
        t = s.rstrip(b" ")
        assert L.ora_hashbpchar_text(t, len(t)) == ref.ref_hash_datum(TEXT, 0, t, len(t)), t
    for ch in b"ANRFO":
        assert L.ora_hash_datum(P.BPCHAR1, ch) == ref.ref_hash_datum(BPCHAR, 0, bytes([ch]), 1)


def test_jump_consistent_hash(oracle, ref):
    """cdbhashreduce's REDUCE_JUMP_HASH on a preset hash value == the oracle's route for a key with that hash"""
    L = oracle.lib()
    rng = random.Random(5)
    t = (C.c_int32 * 1)(P.INT4)
    for _ in range(3000):
        k = rng.getrandbits(32) - 2 ** 31
        nseg = rng.choice([1, 2, 3, 4, 5, 7, 8, 16, 24, 64, 100, 1000])
        h = ref.ref_hash_datum(INT4, k, None, 0)
        want = ref.ref_jump_consistent_hash(h, nseg)
        assert 0 <= want < nseg
        assert L.ora_cdbhash_segment(t, (C.c_int64 * 1)(k), None, 1, nseg) == want


def test_cdbhash_multi_key_with_nulls(oracle, ref):
    """evalHashKey over 1-4 keys of mixed types, NULLs included (a NULL key only rotates, cdbhash.c:196-216)"""
    L = oracle.lib()
    rng = random.Random(23)
    kinds = [(P.INT4, INT4), (P.INT8, INT8), (P.DATE, INT4), (P.FLOAT8, FLOAT8), (P.BPCHAR1, BPCHAR)]
    for _ in range(4000):
        nk = rng.randrange(1, 5)
        nseg = rng.choice([1, 2, 3, 4, 8, 13, 64])
        ot, rk, vals, strs, nulls = [], [], [], [], []
        for _k in range(nk):
            o, r = rng.choice(kinds)
            ot.append(o)
            rk.append(r)
            nulls.append(1 if rng.random() < 0.15 else 0)
            if r == INT4:
                v = rng.getrandbits(32) - 2 ** 31
            elif r == INT8:
                v = rng.getrandbits(64) - 2 ** 63
            elif r == FLOAT8:
                v = _bits(rng.uniform(-1e6, 1e6))
            else:
                v = rng.choice(b"ANRFOxyz")
            vals.append(v)
            strs.append(bytes([v]) if r == BPCHAR else None)
        a = L.ora_cdbhash_segment((C.c_int32 * nk)(*ot), (C.c_int64 * nk)(*vals), (C.c_uint8 * nk)(*nulls), nk, nseg)
        b = ref.ref_cdbhash_segment(nk, (C.c_int * nk)(*rk), (C.c_int64 * nk)(*vals), (C.c_char_p * nk)(*strs),
                                    (C.c_uint8 * nk)(*nulls), nseg)
        assert b >= 0, ref.ref_exec_last_error()
        assert a == b, (ot, vals, nulls, nseg)


def _dec(v, ds):
    """scaled integer -> decimal text with ds fraction digits"""
    s = "-" if v < 0 else ""
    d = str(abs(v)).rjust(ds + 1, "0")
    return s + (d[:-ds] + "." + d[-ds:] if ds else d)


def _split128(v):
    u = v & (2 ** 128 - 1)
    lo, hi = u & (2 ** 64 - 1), u >> 64
    return (lo - 2 ** 64 if lo >= 2 ** 63 else lo), (hi - 2 ** 64 if hi >= 2 ** 63 else hi)


def _both_texts(oracle, v, ds, n):
    E = capi.ex()
    O = oracle.lib()
    lo, hi = _split128(v)
    out = []
    for lib, pre in ((E, "cb"), (O, "ora")):
        a, b = C.create_string_buffer(256), C.create_string_buffer(256)
        getattr(lib, pre + "_numeric_sum_text")(lo, hi, ds, a, 256)
        getattr(lib, pre + "_numeric_avg_text")(lo, hi, ds, n, b, 256)
        out.append((a.value.decode(), b.value.decode()))
    return out


def test_numeric_sum_avg_against_reference_aggregates(oracle, ref):
    """sum / avg(numeric): the reference's numeric_avg_accum + numeric_sum / numeric_avg over the rows == the product's and
    the oracle's finalisation of the exact (sum, N) pair (select_div_scale, round half away from zero)"""
    rng = random.Random(99)
    s_out, a_out = C.create_string_buffer(512), C.create_string_buffer(512)
    cases = 0
    for trial in range(600):
        ds = rng.choice([0, 2, 2, 2, 4, 6])
        n = rng.choice([1, 1, 2, 3, 7, 10, 100, 997])
        mag = rng.choice([1, 10, 10 ** 3, 10 ** 6, 10 ** 9, 10 ** 12, 10 ** 15, 10 ** 17])
        signed = rng.random() < 0.3
        vals = [rng.randrange(-mag if signed else 0, mag + 1) for _ in range(n)]
        if trial % 7 == 0:
            # quotients that round: ...5 ties and long 9 runs
            vals = [rng.choice([1, 2, 5, 10, 125, 9999, 99999999, 2 ** 40 + 1]) for _ in range(n)]
        if trial % 11 == 0:
            vals = [0] * n
        texts = [_dec(v, ds).encode() for v in vals]
        rc = ref.ref_numeric_agg((C.c_char_p * n)(*texts), n, s_out, a_out, 512)
        assert rc == 0, ref.ref_exec_last_error()
        for got in _both_texts(oracle, sum(vals), ds, n):
            assert got == (s_out.value.decode(), a_out.value.decode()), (vals[:5], ds, n)
        cases += 1
    assert cases == 600


def test_numeric_large_sums_against_reference(oracle, ref):
    """sums beyond 64 bits (the 128-bit accumulator): N large rows"""
    rng = random.Random(3)
    s_out, a_out = C.create_string_buffer(512), C.create_string_buffer(512)
    for _ in range(60):
        ds = rng.choice([2, 4, 6])
        n = rng.choice([50, 200])
        vals = [rng.randrange(2 ** 61, 2 ** 63 - 1) * rng.choice([1, 1, -1]) for _ in range(n)]
        if rng.random() < 0.5:
            vals = [abs(v) for v in vals]
        texts = [_dec(v, ds).encode() for v in vals]
        assert ref.ref_numeric_agg((C.c_char_p * n)(*texts), n, s_out, a_out, 512) == 0, ref.ref_exec_last_error()
        for got in _both_texts(oracle, sum(vals), ds, n):
            assert got == (s_out.value.decode(), a_out.value.decode()), (sum(vals), ds, n)


def test_int8_sum_avg_against_reference(oracle, ref):
    """sum / avg(bigint): int8_avg_accum (Int128AggState) + numeric_poly_sum / numeric_poly_avg == finalisation at dscale 0"""
    rng = random.Random(17)
    s_out, a_out = C.create_string_buffer(512), C.create_string_buffer(512)
    for _ in range(300):
        n = rng.choice([1, 2, 3, 10, 64, 500])
        mag = rng.choice([10, 10 ** 4, 10 ** 9, 2 ** 62])
        vals = [rng.randrange(-mag, mag + 1) for _ in range(n)]
        assert ref.ref_int8_agg((C.c_int64 * n)(*vals), n, s_out, a_out, 512) == 0, ref.ref_exec_last_error()
        for got in _both_texts(oracle, sum(vals), 0, n):
            assert got == (s_out.value.decode(), a_out.value.decode()), (vals[:5], n)
    # no rows: both final functions return NULL (the executor's n == 0 branch, cb_exec.c finalize_state)
    assert ref.ref_int8_agg((C.c_int64 * 1)(0), 0, s_out, a_out, 512) == 0
    assert s_out.value == b"" and a_out.value == b""


def test_int4_sum_is_bigint(ref):
    rng = random.Random(1)
    out = C.c_int64()
    for _ in range(100):
        n = rng.choice([1, 5, 1000])
        vals = [rng.randrange(-2 ** 31, 2 ** 31) for _ in range(n)]
        assert ref.ref_int4_sum((C.c_int32 * n)(*vals), n, C.byref(out)) == 0
        assert out.value == sum(vals)
    assert ref.ref_int4_sum((C.c_int32 * 1)(0), 0, C.byref(out)) == 1     # NULL over no rows


def test_scaled_integer_arithmetic_is_numeric_arithmetic(ref):
    """Q1's sum_disc_price / sum_charge arguments: numeric_sub and numeric_mul on numeric(15,2) inputs give exactly the
    int64 product at display scale 4 / 6 (numeric_mul adds dscales, numeric.c:2645), which is what the kernels compute"""
    rng = random.Random(8)
    buf, buf2 = C.create_string_buffer(256), C.create_string_buffer(256)
    for _ in range(2000):
        ext = rng.randrange(90000, 10500000)        # 900.00 .. 104999.99
        disc = rng.randrange(0, 11)                  # 0.00 .. 0.10
        tax = rng.randrange(0, 9)                    # 0.00 .. 0.08
        assert ref.ref_numeric_binop(1, b"1", _dec(disc, 2).encode(), buf, 256) == 0
        assert buf.value.decode() == _dec(100 - disc, 2)
        assert ref.ref_numeric_binop(2, _dec(ext, 2).encode(), buf.value, buf2, 256) == 0
        assert buf2.value.decode() == _dec(ext * (100 - disc), 4)
        assert ref.ref_numeric_binop(0, b"1", _dec(tax, 2).encode(), buf, 256) == 0
        assert ref.ref_numeric_binop(2, buf2.value, buf.value, buf2, 256) == 0
        assert buf2.value.decode() == _dec(ext * (100 - disc) * (100 + tax), 6)


def test_q1_through_reference_functions_matches_expected_rows(oracle, ref, golden):
    """oracle/ref_q1.c: the reference's block writer + reader, numeric_sub / _mul / _add, hashbpchar / bpchareq, numeric_avg_accum,
    numeric_sum / numeric_avg over lineitem reproduce rpt_tpch's expected Q1 rows (rpt_tpch.source:334-340)"""
    from cloudberry_b200 import tpch
    rels, exp = golden
    q = oracle.RefQ1(rels[0])
    rows, passed = q.run(tpch.Q1_CUTOFF)
    assert rows == exp["q1"]
    assert passed == sum(int(r[-1]) for r in rows)
    # a cutoff before every row: no groups (a grouped aggregate over no rows returns no rows)
    assert q.run(-100000) == ([], 0)
    q.free()


@pytest.mark.parametrize("checksum,blocksize", [(True, 32768), (False, 8192), (True, 1 << 20)])
def test_q1_reference_functions_agree_with_oracle_on_synthetic(oracle, ref, checksum, blocksize):
    """the same at a size the reference has no fixture for: 300 000 synthetic lineitem rows, oracle (int64 arithmetic) ==
    reference functions (varlena numeric arithmetic), text for text"""
    from cloudberry_b200 import tpch
    sz = tpch.sizes(1)
    li = tpch._rel("lineitem", tpch.gen_lineitem(42, sz["lineitem"], sz["supplier"], sz["part"], lo=0, hi=300000))
    want = tpch.format_q1(oracle.execute(tpch.q1_plan(1), [[li]], nthreads=1).rows)
    q = oracle.RefQ1(li, checksum=checksum, blocksize=blocksize)
    rows, passed = q.run(tpch.Q1_CUTOFF)
    q.free()
    assert rows == want
    assert len(rows) == 4 and passed == sum(int(r[-1]) for r in rows)


def test_product_hashbpchar_is_the_reference_function(ref):
    """cbgpu_hashbpchar (libcbgpu.so, host side: the per-code hash table of dictionary columns and string constants) ==
    the reference's hashbpchar on random strings, blank padding included"""
    rng = random.Random(41)
    for n in list(range(0, 40)) + [64, 200]:
        s = bytes(rng.choice(b"ABCDEFGHIJKLMNOPQRSTUVWXYZ abcdefghij0123456789") for _ in range(n))
        s += b" " * rng.randrange(0, 6)
        assert capi.hashbpchar(s) == ref.ref_hash_datum(BPCHAR, 0, s, len(s)), s
    for name in ("BUILDING", "AUTOMOBILE", "MACHINERY", "HOUSEHOLD", "FURNITURE", "ASIA", "AMERICA", "UNITED STATES"):
        padded = name.ljust(25).encode()
        assert capi.hashbpchar(name) == ref.ref_hash_datum(BPCHAR, 0, padded, len(padded))


def test_avg_at_full_scale_against_numeric_div(oracle, ref):
    """SF100-sized states: N up to 10^9 rows per group and scaled sums up to 10^24 (beyond 64 bits: the device's 128-bit
    accumulator) - numeric_avg is numeric_div(sumX, N) (numeric.c:6056-6088), so the reference's numeric_div on the texts is
    the expected answer for both finalisers"""
    rng = random.Random(2024)
    buf = C.create_string_buffer(256)
    for _ in range(1500):
        ds = rng.choice([0, 2, 4, 6])
        n = rng.choice([1, 7, 148_000_000, 600_037_902, 999_999_999, rng.randrange(1, 10 ** 9)])
        total = rng.randrange(0, 10 ** rng.choice([3, 9, 15, 18, 19, 20, 22, 24])) * rng.choice([1, 1, 1, -1])
        assert ref.ref_numeric_binop(3, _dec(total, ds).encode(), str(n).encode(), buf, 256) == 0, ref.ref_exec_last_error()
        want_avg = buf.value.decode()
        for got_sum, got_avg in _both_texts(oracle, total, ds, n):
            assert got_sum == _dec(total, ds)
            assert got_avg == want_avg, (total, ds, n)


def test_avg_at_the_edges_of_the_128_bit_state(oracle, ref):
    """sums at +-2^127, N up to 2^63 - 1, display scales 0-12, powers of NBASE as divisors (the weight / first-digit rule of
    select_div_scale flips there)"""
    rng = random.Random(5)
    buf = C.create_string_buffer(512)
    for _ in range(4000):
        ds = rng.choice([0, 1, 2, 3, 4, 6, 9, 12])
        v = rng.choice([2 ** 127 - 1, -2 ** 127, rng.randrange(-10 ** 38, 10 ** 38), rng.randrange(-10 ** 30, 10 ** 30),
                        rng.randrange(-10 ** 12, 10 ** 12), rng.randrange(0, 10 ** 5)])
        n = rng.choice([1, 3, 9, 9999, 10000, 10001, 2 ** 31, 2 ** 62, 2 ** 63 - 1, rng.randrange(1, 10 ** 15), 10 ** rng.randrange(0, 18)])
        assert ref.ref_numeric_binop(3, _dec(v, ds).encode(), str(n).encode(), buf, 512) == 0, ref.ref_exec_last_error()
        for got_sum, got_avg in _both_texts(oracle, v, ds, n):
            assert got_sum == _dec(v, ds)
            assert got_avg == buf.value.decode(), (v, ds, n)
