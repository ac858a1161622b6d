"""Oracle- and golden-anchored parity ABOVE toy size (-m gpu): the specialised kernels AND the generic interpreter against
  * the CPU oracle at SF1 (6 M lineitem rows: seconds on the host), Q1 / Q3 / Q5 and SSB Q4.x;
  * the independent numpy answers at SF10 (60 M rows; tests/golden/bench_golden.json, tools/make_bench_golden.py --only sf10);
and the limit that queue entries are 32-bit row ids: a relation beyond 2^32 - 16 rows is refused loudly, not truncated."""
import pytest

from cloudberry_b200 import bench_golden as BG
from cloudberry_b200 import capi, harness, ssb, tpch
from gpu_util import to_device

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("generic", [False, True])
def test_sf1_against_the_oracle(ctx, oracle, generic):
    rels_o = tpch.gen_tables(1, oracle.hashbpchar)
    dev, _ = harness.device_tables(ctx, 1)              # device generator: the same formulas (checked here against the host's)
    ex = capi.Executor(ctx, dev, force_generic=generic)
    seg, reg = tpch.SEGMENTS.index("MACHINERY"), tpch.REGIONS.index("AMERICA")
    for plan, fmt in ((tpch.q1_plan(1), tpch.format_q1), (tpch.q3_plan(seg, 1), tpch.format_q3),
                      (tpch.q5_plan(reg, 1), lambda r: tpch.format_q5(r, tpch.NATIONS))):
        assert fmt(ex.run(plan).rows) == fmt(oracle.execute(plan, [rels_o]).rows)
    ex.close()
    for d in dev:
        d.free()
    srels_o = ssb.gen_tables(1, oracle.hashbpchar)
    sdev, _ = ssb.device_tables(ctx, 1, capi.hashbpchar)
    exs = capi.Executor(ctx, sdev, force_generic=generic)
    for q in ("q4.1", "q4.2", "q4.3"):
        plan = ssb.PLANS[q]()
        assert ssb.canon(exs.run(plan).rows) == ssb.canon(oracle.execute(plan, [srels_o]).rows)
    exs.close()
    for d in sdev:
        d.free()


@pytest.mark.parametrize("generic", [False, True])
def test_sf10_against_the_numpy_goldens(ctx, generic):
    gold = BG.load()
    dev, sz = harness.device_tables(ctx, 10)
    assert sz["lineitem"] == gold["q1_sf10"]["rows_per_shard"]
    ex = capi.Executor(ctx, dev, force_generic=generic)
    assert tpch.format_q1(ex.run(tpch.q1_plan(1)).rows) == BG.q1_rows(gold["q1_sf10"], 1)
    assert tpch.format_q3(ex.run(tpch.q3_plan(tpch.SEGMENTS.index("MACHINERY"), 1)).rows) == BG.q3_rows(gold["q3_sf10"])
    assert tpch.format_q5(ex.run(tpch.q5_plan(tpch.REGIONS.index("AMERICA"), 1)).rows, tpch.NATIONS) == BG.q5_rows(gold["q5_sf10"])
    ex.close()
    for d in dev:
        d.free()
    sdev, _ = ssb.device_tables(ctx, 10, capi.hashbpchar)
    exs = capi.Executor(ctx, sdev, force_generic=generic)
    for q in ("q4.1", "q4.2", "q4.3"):
        assert ssb.canon(exs.run(ssb.PLANS[q]()).rows) == BG.ssb_rows(gold["ssb_sf10"], q)
    exs.close()
    for d in sdev:
        d.free()


def test_relation_beyond_32_bit_row_ids_is_refused(ctx):
    """row ids travel as uint32 (queue entries, hash table slots): a relation the kernels could not address is refused when it is
    created (cbgpu_rel_create) - loudly, with the limit in the message - instead of being truncated later"""
    from cloudberry_b200 import plan as P
    with pytest.raises(capi.CbgpuError) as e:
        capi.DeviceRelation(ctx, (1 << 32) - 8, [P.BPCHAR1], name="huge")
    assert "row" in str(e.value).lower()
    ok = capi.DeviceRelation(ctx, 1 << 20, [P.BPCHAR1], name="fine")
    ok.free()
