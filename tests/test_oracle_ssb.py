"""SSB Q4.x (BASELINE.json configs[4]) on the CPU oracle (not gpu).

The reference has no SSB fixture, so the oracle is checked against an independent numpy
evaluation of the same SQL over the same seeded tables (cloudberry_b200/ssb.py numpy_answer).
"""
import pytest

from cloudberry_b200 import ssb


@pytest.mark.parametrize("q", ["q4.1", "q4.2", "q4.3"])
def test_oracle_matches_numpy(oracle, q):
    rels = ssb.gen_tables(0.01, oracle.hashbpchar)
    res = oracle.execute(ssb.PLANS[q](), [rels])
    want = ssb.numpy_answer(q, rels)
    assert len(want) > 0
    assert ssb.canon(res.rows) == want
