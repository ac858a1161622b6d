"""CPU-side checks of the product (not gpu): the C-ABI libraries load and export every symbol the
headers declare, and the host-side logic (numeric finalisation, dictionary hashing) agrees with the
oracle.  No kernel is launched here."""
import ctypes as C
import random

from cloudberry_b200 import capi


def test_libraries_export_every_declared_symbol():
    G = capi.gpu()
    E = capi.ex()
    missing = []
    for name in capi.header_symbols("cbgpu.h"):
        if not hasattr(G, name):
            missing.append(name)
    for name in capi.header_symbols("cb_exec.h"):
        if name.startswith("cbgpu_"):
            continue
        if not hasattr(E, name):
            missing.append(name)
    assert not missing, missing
    assert len(capi.header_symbols("cbgpu.h")) > 40


def test_numeric_finalisation_matches_oracle(oracle):
    """product: base-1e9 limb arithmetic (csrc/exec/cb_numeric.c); oracle: digit strings."""
    E = capi.ex()
    O = oracle.lib()
    rng = random.Random(11)
    a = C.create_string_buffer(200)
    b = C.create_string_buffer(200)
    cases = [(38045600, 2, 14876), (53234821165, 2, 14876), (74501, 2, 14876), (0, 2, 5), (1, 0, 3), (2, 0, 3),
             (10, 0, 4), (999999, 4, 7), (5, 6, 1000000007), (-38045600, 2, 14876), (10 ** 30, 6, 600000000),
             (123456789012345678901234567890, 6, 1), (1, 6, 9223372036854775807), (49, 2, 98), (50, 2, 100),
             (9999, 0, 10000), (10000, 0, 10000), (99995, 4, 2)]
    for _ in range(3000):
        mag = rng.choice([3, 8, 15, 22, 30, 37])
        s = rng.randrange(-10 ** mag, 10 ** mag)
        cases.append((s, rng.choice([0, 2, 4, 6]), rng.choice([1, 2, 3, 7, 348, 14876, 10 ** 6, 6 * 10 ** 8, 2 ** 40 + 1])))
    for s, ds, n in cases:
        lo = s & (2 ** 64 - 1)
        lo = lo - 2 ** 64 if lo >= 2 ** 63 else lo
        hi = (s >> 64)
        E.cb_numeric_sum_text(lo, hi, ds, a, 200)
        O.ora_numeric_sum_text(lo, hi, ds, b, 200)
        assert a.value == b.value, (s, ds)
        E.cb_numeric_avg_text(lo, hi, ds, n, a, 200)
        O.ora_numeric_avg_text(lo, hi, ds, n, b, 200)
        assert a.value == b.value, (s, ds, n, a.value, b.value)


def test_numeric_avg_against_python_decimal():
    """Independent third opinion: Python's decimal with ROUND_HALF_UP at the reference's scale."""
    from decimal import Decimal, ROUND_HALF_UP, getcontext
    getcontext().prec = 200
    E = capi.ex()
    a = C.create_string_buffer(200)
    for s, ds, n in [(38045600, 2, 14876), (53234821165, 2, 14876), (74501, 2, 14876), (1016280268835844, 6, 28818)]:
        E.cb_numeric_avg_text(s, 0, ds, n, a, 200)
        txt = a.value.decode()
        rscale = len(txt.split(".")[1])
        want = (Decimal(s) / (Decimal(10) ** ds) / Decimal(n)).quantize(Decimal(1).scaleb(-rscale), rounding=ROUND_HALF_UP)
        assert Decimal(txt) == want


def test_hashbpchar_host_matches_oracle(oracle):
    for t in ["", "A", "MACHINERY", "UNITED STATES", "MIDDLE EAST     ", "x" * 40, "1-URGENT", " "]:
        assert capi.hashbpchar(t) == oracle.hashbpchar(t)


def test_no_device_means_a_loud_error_not_a_fallback():
    """On a box without a usable GPU every way into the product ends in an error that says so: the context cannot be
    created (CbgpuError with the CUDA runtime's message), and the operator entry points refuse a NULL context / executor
    state instead of computing anything on the host.  (Skipped where a device is visible: there the -m gpu tests apply.)"""
    import ctypes as C

    import pytest
    G = capi.gpu()
    if G.cbgpu_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(capi.CbgpuError) as ei:
        capi.Context(0)
    assert ei.value.code != 0 and "cuda" in str(ei.value).lower()
    E = capi.ex()
    # no context -> no executor state; no executor state -> no plan state (execProcnode.c:190's contract: NULL in, NULL out)
    es = E.cb_CreateExecutorState(None, None, 0)
    if es:
        from cloudberry_b200 import tpch
        from cloudberry_b200 import plan as P
        ps = E.cb_ExecInitNode(P.plan_ptr(tpch.q1_plan(1)), es, 0)
        if ps:
            # initialisation is lazy (streams are opened by the first pull): the first ExecProcNode must fail, with a message
            slot = E.cb_ExecProcNode(ps)
            assert (not slot) or slot.contents.tts_empty
            E.cb_ExecEndNode(ps)
        assert es.contents.es_errcode != 0
        assert E.cb_estate_error(es)
        E.cb_FreeExecutorState(es)
