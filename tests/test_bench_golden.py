"""tests/golden/bench_golden.json (the full-size answers bench.py checks its rows against) and the code that made it.

The full-size file cannot be re-derived here (12 minutes of all cores), but the SAME code can be run at a size the CPU oracle
follows: tools/make_bench_golden.py's arithmetic evaluation of Q1 / Q3 / Q5 / SSB Q4.x and cloudberry_b200/bench_golden.py's
text formatting (numeric display scales, avg rounding) must agree with the oracle's executor row for row."""
import json
import multiprocessing as mp
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from cloudberry_b200 import bench_golden as BG  # noqa: E402
from cloudberry_b200 import ssb, tpch  # noqa: E402


def test_full_size_file_is_complete():
    g = json.load(open(BG.PATH))
    assert len(g["q1_sf100"]["shards"]) == 8 and g["q1_sf100"]["rows_per_shard"] == 600037902
    for sh in g["q1_sf100"]["shards"]:
        assert sorted(sh) == ["AF", "NF", "NO", "RF"]
    assert len(g["q3_sf100"]["rows"]) == 10 and g["q3_sf100"]["groups"] > 1000000
    assert len(g["q5_sf100"]["rows"]) == 5 and len(g["q5_sf300"]["rows"]) == 5
    assert [len(g["ssb_sf100"][q]) for q in ("q4.1", "q4.2", "q4.3")] == [35, 100, 800]
    # Q1's counts over the 8 shards must add up to an independent count: every row is in exactly one group or fails the qual
    total = sum(int(st[0]) for sh in g["q1_sf100"]["shards"] for st in sh.values())
    assert 0.97 * 8 * 600037902 < total < 8 * 600037902


@pytest.mark.parametrize("sf", [0.05, 0.3])
def test_generator_of_the_golden_file_agrees_with_the_oracle(oracle, sf):
    import make_bench_golden as G
    with mp.Pool(4) as pool:
        g1, g3, g5, gs = G.run_q1(pool, sf, 2), G.run_q3(pool, sf), G.run_q5(pool, sf), G.run_ssb(pool, sf)
    sz = tpch.sizes(sf)
    # two Q1 shards = generator rows [0, 2 n)
    li = tpch._rel("lineitem", tpch.gen_lineitem(42, sz["lineitem"], sz["supplier"], sz["part"], lo=0, hi=2 * sz["lineitem"]))
    assert tpch.format_q1(oracle.execute(tpch.q1_plan(1), [[li]]).rows) == BG.q1_rows(g1, 2)
    rels = tpch.gen_tables(sf, oracle.hashbpchar)
    assert tpch.format_q3(oracle.execute(tpch.q3_plan(tpch.SEGMENTS.index("MACHINERY"), 1), [rels]).rows) == BG.q3_rows(g3)
    assert tpch.format_q5(oracle.execute(tpch.q5_plan(tpch.REGIONS.index("AMERICA"), 1), [rels]).rows, tpch.NATIONS) == BG.q5_rows(g5)
    srels = ssb.gen_tables(sf, oracle.hashbpchar)
    for q in ("q4.1", "q4.2", "q4.3"):
        assert ssb.canon(oracle.execute(ssb.PLANS[q](), [srels]).rows) == BG.ssb_rows(gs, q) == ssb.numpy_answer(q, srels)


def test_avg_text_follows_select_div_scale():
    # the golden Q1 row of the reference's regression test (rpt_tpch.source:334-340), group A / F
    assert BG.avg_text(38045600, 2, 14876) == "25.5751546114546921"
    assert BG.avg_text(53234821165, 2, 14876) == "35785.709306937349"
    assert BG.scaled_text(-5, 2) == "-0.05" and BG.scaled_text(123456, 4) == "12.3456"
    assert BG.check("q", [[1]], [[1]]) == "ok" and BG.check("q", [[1]], [[2]]).startswith("MISMATCH")
