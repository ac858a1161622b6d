"""Worker of tests/test_host_logic_fake_runtime.py: drives the host executor over tests/native/fake_cudart.c (LD_PRELOADed by
the test; kernels are no-ops, "device" memory is zeroed host memory).  Prints one JSON object.  Not a test by itself."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from cloudberry_b200 import capi, ssb, tpch  # noqa: E402
from cloudberry_b200 import plan as P  # noqa: E402

if os.environ.get("CB_TEST_LIBDIR"):
    capi.HERE = os.environ["CB_TEST_LIBDIR"]        # libcbgpu.so (link) + a sanitizer build of libcbexec.so


def main():
    out = {"ran": {}, "refused": {}}
    G = capi.gpu()
    assert G.cbgpu_device_count() == 1
    ctx = capi.Context(0)
    rels, _ = tpch.load_golden(capi.hashbpchar)
    dev = [capi.DeviceRelation.from_host(ctx, r) for r in rels]
    seg = tpch.SEGMENTS.index("MACHINERY")
    reg = tpch.REGIONS.index("AMERICA")

    # 1. every plan shape of the metric goes through translation, program emission, kernel matching, launch and read-back
    for generic in (False, True):
        ex = capi.Executor(ctx, dev, force_generic=generic)
        for name, plan in (("q1", tpch.q1_plan(1)), ("q3", tpch.q3_plan(seg, 1)), ("q5", tpch.q5_plan(reg, 1))):
            before = ctx.launches()
            res = ex.run(plan)
            out["ran"]["%s%s" % (name, "_generic" if generic else "")] = {"rows": len(res.rows), "launches": ctx.launches() - before}
        ex.close()

    # 2. plans with Motions on a 3-segment cluster in one process (two-stage aggregation, redistributed joins)
    cl = capi.Cluster(ctx, [dev, dev, dev])
    for name, plan in (("q1_3seg", tpch.q1_plan(3)), ("q3_3seg", tpch.q3_plan(seg, 3, customer_replicated=False)),
                       ("q5_3seg", tpch.q5_plan(reg, 3, replicated=False))):
        before = ctx.launches()
        res = cl.run(plan)
        out["ran"][name] = {"rows": len(res.rows), "launches": ctx.launches() - before}
    cl.close()

    # 3. what the host must refuse, loudly and before any kernel: (plan builder, expected message fragment)
    li = rels[0]
    sc = P.SeqScan(1, [(n, P.Var(1, li.attno(n), *li.var(n)[1:])) for n in ("l_returnflag", "l_quantity", "l_orderkey")])
    v = tpch._child_var(sc)
    keys_t = [("l_returnflag", v("l_returnflag")), ("s", P.Aggref(P.AGG_SUM, v("l_quantity")))]

    def having():
        return P.Agg(sc, P.AGG_HASHED, P.AGGSPLIT_SIMPLE, [1], keys_t, num_groups=8,
                     quals=[P.OpExpr(P.OP_GT, P.OuterVar(2, P.NUMERIC, 2), P.NumericConst("1"))])

    def sorted_agg():
        return P.Agg(sc, P.AGG_SORTED, P.AGGSPLIT_SIMPLE, [1], keys_t, num_groups=8)

    def sort_without_limit():
        a = P.Agg(sc, P.AGG_HASHED, P.AGGSPLIT_SIMPLE, [1], keys_t, num_groups=8)
        return P.LimitSort(a, [(1, False)], -1)

    def right_join():
        o = rels[1]
        so = P.SeqScan(2, [("o_orderkey", P.Var(2, o.attno("o_orderkey"), P.INT8))])
        h = P.Hash(so, [P.out_var(so, 1)])
        j = P.HashJoin(P.JOIN_RIGHT, sc, h, [P.out_var(sc, 3)], [("l_quantity", P.out_var(sc, 2))])
        return P.Agg(j, P.AGG_PLAIN, P.AGGSPLIT_SIMPLE, [], [("n", P.Aggref(P.AGG_COUNT_STAR))])

    def numeric_join_key():
        s2 = P.SeqScan(1, [("l_quantity", P.Var(1, li.attno("l_quantity"), P.NUMERIC, 2))])
        h = P.Hash(s2, [P.out_var(s2, 1)])
        j = P.HashJoin(P.JOIN_INNER, sc, h, [P.out_var(sc, 2)], [("l_quantity", P.out_var(sc, 2))])
        return P.Agg(j, P.AGG_PLAIN, P.AGGSPLIT_SIMPLE, [], [("n", P.Aggref(P.AGG_COUNT_STAR))])

    def bad_scanrelid():
        s9 = P.SeqScan(9, [("x", P.Var(9, 1, P.INT8))])
        return P.Agg(s9, P.AGG_PLAIN, P.AGGSPLIT_SIMPLE, [], [("n", P.Aggref(P.AGG_COUNT_STAR))])

    ex = capi.Executor(ctx, dev)
    for name, build in (("having", having), ("sorted_agg", sorted_agg), ("sort_without_limit", sort_without_limit),
                        ("right_join", right_join), ("numeric_join_key", numeric_join_key), ("bad_scanrelid", bad_scanrelid)):
        before = ctx.launches()
        try:
            ex.run(build())
            out["refused"][name] = {"error": None}
        except capi.CbgpuError as e:
            out["refused"][name] = {"error": str(e), "code": e.code, "launches": ctx.launches() - before}
    # the executor state is reusable after a refusal
    out["ran"]["q1_after_refusals"] = {"rows": len(ex.run(tpch.q1_plan(1)).rows)}
    ex.close()

    # 4. SSB Q4.x: wide group-by plans
    srels = ssb.gen_tables(0.01, capi.hashbpchar)
    if srels is not None:
        sdev = [capi.DeviceRelation.from_host(ctx, r) for r in srels]
        ex = capi.Executor(ctx, sdev)
        for q in ("q4_1", "q4_2", "q4_3"):
            out["ran"]["ssb_" + q] = {"rows": len(ex.run(getattr(ssb, q + "_plan")()).rows)}
        ex.close()
        for d in sdev:
            d.free()
    for d in dev:
        d.free()
    ctx.close()
    print("HOSTLOGIC " + json.dumps(out))


if __name__ == "__main__":
    main()
