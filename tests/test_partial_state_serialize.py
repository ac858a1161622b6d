"""Partial aggregate states in the reference's wire form: what a device Partial Aggregate hands to a
CPU Finalize stage, and back.  csrc/exec/cb_numeric.c (cb_numeric_avg_serialize, cb_int8_avg_serialize,
cb_numeric_avg_deserialize) against the reference's OWN serialisation / deserialisation / final functions
(numeric_avg_serialize :5025, numeric_avg_deserialize :5092, int8_avg_serialize :5793, int8_avg_deserialize, numeric_sum / _avg,
numeric_poly_sum / _avg of utils/adt/numeric.c, with libpq/pqformat.c and common/stringinfo.c, compiled where they lie into
oracle/_ref/libexec_ref.so):

  * the bytes we make of (N, exact sum) are the bytes the reference makes after accumulating the same inputs - bit for bit;
  * the reference's Finalize over OUR bytes prints the sum / avg it prints over its own;
  * our deserialiser reads the reference's bytes back into the same (N, sum, scale)."""
import ctypes as C
import os
import random
from decimal import Decimal

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFLIB = os.path.join(ROOT, "oracle", "_ref", "libexec_ref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REFLIB), reason="oracle/_ref/libexec_ref.so is built from /root/reference (make -C oracle ref)")


@pytest.fixture(scope="module")
def libs():
    R = C.CDLL(REFLIB)
    E = C.CDLL(os.path.join(ROOT, "cloudberry_b200", "libcbexec.so"))
    E.cb_numeric_avg_serialize.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int32]
    E.cb_int8_avg_serialize.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int32]
    E.cb_numeric_avg_deserialize.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                             C.POINTER(C.c_int32)]
    return R, E


def _split(v):
    """a Python int as the (lo, hi) signed 64-bit halves of a 128-bit two's complement value"""
    u = v & ((1 << 128) - 1)
    lo, hi = u & ((1 << 64) - 1), u >> 64
    return lo - (1 << 64) if lo >= 1 << 63 else lo, hi - (1 << 64) if hi >= 1 << 63 else hi


def _join(lo, hi):
    return (hi << 64) | (lo & ((1 << 64) - 1))


CASES = [
    ("q1 prices", 2, lambda r: r.randrange(90000, 10500000), 5000),
    ("small positive", 2, lambda r: r.randrange(0, 11), 300),
    ("mixed sign", 4, lambda r: r.randrange(-10**9, 10**9), 2000),
    ("zeros only", 2, lambda r: 0, 50),
    ("one value", 0, lambda r: r.randrange(-5, 6), 1),
    ("near 64 bits", 2, lambda r: r.randrange(2**62, 2**63 - 1), 4000),          # the sum leaves 64 bits
    ("negative near 64 bits", 6, lambda r: -r.randrange(2**62, 2**63 - 1), 3000),
    ("trailing zero digits", 3, lambda r: 1000 * r.randrange(1, 10**6), 777),
    ("scale 9", 9, lambda r: r.randrange(-10**12, 10**12), 1234),
]


@pytest.mark.parametrize("name,dscale,gen,count", CASES, ids=[c[0] for c in CASES])
def test_numeric_states_match_the_reference(libs, name, dscale, gen, count):
    R, E = libs
    rnd = random.Random(hash(name) & 0xffff)
    scaled = [gen(rnd) for _ in range(count)]
    # numeric literals must carry the column's display scale (as values of a numeric(p, s) column do)
    texts = [format(Decimal(v).scaleb(-dscale), "f") if dscale else str(v) for v in scaled]
    arr = (C.c_char_p * count)(*[t.encode() for t in texts])
    ref = C.create_string_buffer(512)
    nref = R.ref_numeric_avg_serialize(arr, count, ref, 512)
    assert nref > 0
    lo, hi = _split(sum(scaled))
    ours = C.create_string_buffer(512)
    nours = E.cb_numeric_avg_serialize(count, lo, hi, dscale, ours, 512)
    assert nours == nref and ours.raw[:nours] == ref.raw[:nref]
    # the reference's Finalize over our bytes
    s1, a1, s2, a2 = (C.create_string_buffer(256) for _ in range(4))
    assert R.ref_numeric_avg_finalize(ours, nours, s1, a1, 256) == 0
    assert R.ref_numeric_agg(arr, count, s2, a2, 256) == 0
    assert (s1.value, a1.value) == (s2.value, a2.value)
    # and our deserialiser over the reference's bytes
    n, dlo, dhi, ds = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
    assert E.cb_numeric_avg_deserialize(ref, nref, 1, C.byref(n), C.byref(dlo), C.byref(dhi), C.byref(ds)) == 0
    assert n.value == count and ds.value == dscale
    assert _join(dlo.value, dhi.value) == sum(scaled)


@pytest.mark.parametrize("count,lo_v,hi_v", [(1, -7, 8), (1000, -2**62, 2**62), (5000, 2**62, 2**63 - 1), (3000, -(2**63), -(2**62))])
def test_int8_states_match_the_reference(libs, count, lo_v, hi_v):
    R, E = libs
    rnd = random.Random(count)
    vals = [rnd.randrange(lo_v, hi_v) for _ in range(count)]
    arr = (C.c_int64 * count)(*vals)
    ref = C.create_string_buffer(256)
    nref = R.ref_int8_avg_serialize(arr, count, ref, 256)
    assert nref > 0
    lo, hi = _split(sum(vals))
    ours = C.create_string_buffer(256)
    nours = E.cb_int8_avg_serialize(count, lo, hi, ours, 256)
    assert nours == nref and ours.raw[:nours] == ref.raw[:nref]
    s1, a1, s2, a2 = (C.create_string_buffer(256) for _ in range(4))
    assert R.ref_int8_avg_finalize(ours, nours, s1, a1, 256) == 0
    assert R.ref_int8_agg(arr, count, s2, a2, 256) == 0
    assert (s1.value, a1.value) == (s2.value, a2.value)
    n, dlo, dhi, ds = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
    assert E.cb_numeric_avg_deserialize(ref, nref, 0, C.byref(n), C.byref(dlo), C.byref(dhi), C.byref(ds)) == 0
    assert n.value == count and ds.value == 0 and _join(dlo.value, dhi.value) == sum(vals)


def test_unrepresentable_states_are_refused(libs):
    R, E = libs
    n, dlo, dhi, ds = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
    # a NaN went into the CPU partial stage: the reference's bytes carry NaNcount = 1
    arr = (C.c_char_p * 2)(b"1.50", b"NaN")
    ref = C.create_string_buffer(256)
    nref = R.ref_numeric_avg_serialize(arr, 2, ref, 256)
    assert nref > 0
    assert E.cb_numeric_avg_deserialize(ref, nref, 1, C.byref(n), C.byref(dlo), C.byref(dhi), C.byref(ds)) == -2
    # truncated / padded input
    ok = C.create_string_buffer(256)
    k = E.cb_numeric_avg_serialize(3, 12345, 0, 2, ok, 256)
    assert E.cb_numeric_avg_deserialize(ok, k - 1, 1, C.byref(n), C.byref(dlo), C.byref(dhi), C.byref(ds)) == -1
    assert E.cb_numeric_avg_serialize(3, 12345, 0, 2, ok, 10) == -1
