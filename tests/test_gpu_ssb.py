"""SSB Q4.x (BASELINE.json configs[4]: star join of four dimensions, wide hash aggregate) through the
C ABI on the GPU: exact equality with the CPU oracle and with the independent numpy answer."""
import pytest

from cloudberry_b200 import capi, ssb
from gpu_util import to_device

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("q", ["q4.1", "q4.2", "q4.3"])
def test_ssb_q4(ctx, oracle, q, generic):
    rels_o = ssb.gen_tables(0.02, oracle.hashbpchar)
    rels_p = ssb.gen_tables(0.02, capi.hashbpchar)
    dev = to_device(ctx, rels_p)
    ex = capi.Executor(ctx, dev, force_generic=generic)
    plan = ssb.PLANS[q]()
    got = ex.run(plan)
    want = oracle.execute(plan, [rels_o])
    assert ssb.canon(got.rows) == ssb.canon(want.rows) == ssb.numpy_answer(q, rels_o)
    ex.close()
    for d in dev:
        d.free()
