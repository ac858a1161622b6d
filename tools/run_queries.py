#!/usr/bin/env python3
"""Run TPC-H Q1 / Q3 / Q5 (and SSB Q4.x: --queries ssb4.1,ssb4.2,ssb4.3) over device-generated synthetic tables and print timings
(harness / profiling aid; bench.py is the contract benchmark)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cloudberry_b200 import capi, tpch  # noqa: E402


from cloudberry_b200.harness import device_tables  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--queries", default="q1,q3,q5")
    ap.add_argument("--generic", action="store_true")
    ap.add_argument("--trace", action="store_true", help="print every kernel launch of one extra run with its time")
    args = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    ctx = capi.Context(local)
    motion = None
    if world > 1:
        # under torchrun: one SF-sized database distributed over the ranks, plans with Motions over NCCL
        import torch
        import torch.distributed as dist
        from cloudberry_b200 import harness
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ids = [capi.Motion.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        motion = capi.Motion(ctx, rank, world, ids[0])
        rt, sz = harness.distributed_tables(ctx, motion, args.sf, rank, world)
    else:
        rt, sz = device_tables(ctx, args.sf)
    ex = capi.Executor(ctx, rt, force_generic=args.generic, motion=motion)
    ssb_ex = None
    if any(q.startswith("ssb") for q in args.queries.split(",")):
        from cloudberry_b200 import ssb
        ssb_dev, ssb_sz = ssb.device_tables(ctx, args.sf, capi.hashbpchar, rank, world)
        ssb_ex = capi.Executor(ctx, ssb_dev, force_generic=args.generic, motion=motion)
    plans = {"q1": lambda: tpch.q1_plan(world), "q3": lambda: tpch.q3_plan(tpch.SEGMENTS.index("MACHINERY"), world, customer_replicated=False),
             "q5": lambda: tpch.q5_plan(tpch.REGIONS.index("AMERICA"), world, replicated=False)}
    rows_in = {"q1": sz["lineitem"], "q3": sz["lineitem"] + sz["orders"] + sz["customer"],
               "q5": sz["lineitem"] + sz["orders"] + sz["customer"] + sz["supplier"] + 30}
    for q in args.queries.split(","):
        if q.startswith("ssb"):
            plans[q] = lambda q=q: ssb.PLANS["q" + q[3:]](world)
            rows_in[q] = ssb.query_rows_bytes(ssb_sz)[0]
        plan = plans[q]()
        ex_q = ssb_ex if q.startswith("ssb") else ex
        res = ex_q.run(plan)       # warm-up
        times = []
        for _ in range(args.steps):
            ctx.kernel_log_reset()
            ctx.timer_start()
            res = ex_q.run(plan)
            times.append(ctx.timer_stop_ms())
        kn, km = ctx.longest_kernel()
        if args.trace:
            ctx.trace_begin()
            ex_q.run(plan)
            tr = ctx.trace_end()
            if rank != 0:
                continue
            print("trace %s: %d launches, %.3f ms" % (q, len(tr), sum(m for _, m in tr)))
            for name, m in tr:
                print("   %-28s %9.3f ms" % (name, m))
        ms = sorted(times)[len(times) // 2]
        if rank != 0:
            continue
        print(json.dumps({"query": q, "sf": args.sf, "ms": ms, "rows_per_s": rows_in[q] / (ms / 1e3), "result_rows": len(res.rows),
                          "longest_kernel": kn, "longest_kernel_ms": km,
                          "nodes": {k: {"node": v["node"], "kernels": v["kernels"], "device_ms": round(v["device_ms"], 3),
                                        "ntuples": v["ntuples"]} for k, v in res.instrument.items()},
                          "first_rows": res.rows[:3]}))
    ex.close()
    ctx.close()


if __name__ == "__main__":
    main()
