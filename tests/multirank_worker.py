"""Worker for the multi-GPU parity test: run under torchrun, one rank per GPU-segment.

Every rank holds its cdbhash shard of the reference's regression fixture (placement computed by the
oracle, the GPU path never sees the other shards), joins the NCCL interconnect, and runs the
two-stage Q1 and the Motion-bearing Q3 / Q5 plans through the executor.  Rank 0 (the Gather Motion's
receiver) checks the rows against the reference's expected output."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from cloudberry_b200 import capi, tpch
    from oracle import oracle as O
    from gpu_util import shard, to_device

    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    # CB_BOOT=gloo: the interconnect is the peer-memory windows alone, bootstrapped through a gloo all-gather (no NCCL
    # communicator), and all ranks may share ONE device: the single-GPU box runs the inter-process transport this way
    boot_gloo = os.environ.get("CB_BOOT") == "gloo"
    if boot_gloo:
        local = local % capi.gpu().cbgpu_device_count()
        dist.init_process_group("gloo")
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rels, exp = tpch.load_golden(capi.hashbpchar)
    replicated = os.environ.get("CB_REPLICATED", "1") == "1"
    dist_keys = dict(tpch.DIST_KEY)
    if not replicated:
        dist_keys["customer"] = "c_custkey"
        dist_keys["supplier"] = "s_suppkey"
    mine = shard(O, rels, world, dist_keys)[rank]
    ctx = capi.Context(local)
    dev = to_device(ctx, mine)
    if boot_gloo:
        def allgather(mine):
            box = [None] * world
            dist.all_gather_object(box, mine)
            return box
        motion = capi.Motion(ctx, rank, world, allgather=allgather)
    else:
        ids = [capi.Motion.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        motion = capi.Motion(ctx, rank, world, ids[0])
    ex = capi.Executor(ctx, dev, motion=motion)
    seg = exp["dict"]["c_mktsegment_dict"].index("MACHINERY")
    reg = exp["dict"]["r_name_dict"].index("AMERICA")
    ok = True
    # load-time DISTRIBUTED BY: every rank starts from a row-range slice of lineitem, the Redistribute Motion
    # (batch output, cb_ExecProcNodeBatch) must leave it with exactly the oracle's cdbhash shard
    import numpy as np
    from cloudberry_b200 import harness
    li = rels[0]
    lo, hi = rank * li.nrows // world, (rank + 1) * li.nrows // world
    sl = capi.DeviceRelation.from_host(ctx, li.take(np.arange(lo, hi)))
    mine_li = harness.distribute_by_hash(ctx, motion, sl, li.attno("l_orderkey") - 1, "lineitem")
    got = sorted(zip(*[mine_li.read_column(c)[0].tolist() for c in range(len(li.names))]))
    want = sorted(zip(*[c.tolist() for c in mine[0].columns]))
    if got != want:
        ok = False
        print("MULTIRANK FAIL: distribute_by_hash gave %d rows, the oracle's shard has %d" % (len(got), len(want)))
    mine_li.free()
    sl.free()
    # a null map that exists on ONE segment only: the direct Redistribute carries NULL bytes from the senders that have
    # them (the mask rides on the completion signal); the Gather's direct-or-staged decision rides on its signal too
    from cloudberry_b200 import plan as P
    from cloudberry_b200.relation import HostRelation
    kk = np.arange(100, dtype=np.int64) + 1000 * rank
    nl = (np.arange(100) % 10 == 0).astype(np.uint8) if rank == 0 else None
    t = HostRelation("t", ["k", "v"], [P.INT4, P.INT8], [kk, kk * 2], nulls=[nl, None])
    dt = capi.DeviceRelation.from_host(ctx, t)
    ext = capi.Executor(ctx, [dt], motion=motion)
    sc = P.SeqScan(1, [("k", P.Var(1, 1, P.INT4)), ("v", P.Var(1, 2, P.INT8))])
    mh = P.Motion(sc, P.MOTIONTYPE_HASH, [P.out_var(sc, 1)], world)
    rt = ext.run(P.Motion(mh, P.MOTIONTYPE_GATHER))
    if rank == 0:
        want_t = []
        for r in range(world):
            for i in range(100):
                k = i + 1000 * r
                want_t.append((None if (r == 0 and i % 10 == 0) else k, 2 * k))
        got_t = sorted(((-1 if a is None else a), b) for a, b in (tuple(x) for x in rt.rows))
        if got_t != sorted(((-1 if a is None else a), b) for a, b in want_t):
            ok = False
            print("MULTIRANK FAIL: asymmetric null map through Redistribute + Gather: %d rows" % len(rt.rows))
    ext.close()
    dt.free()
    # skew: nine rows of ten go to ONE destination (the reference sends tuple by tuple and cannot overflow,
    # cdbmotion.c:425): the Motion must deliver every row, direct (the window takes it) or by the exact-size staged redo
    nsk = int(os.environ.get("CB_SKEW_ROWS", "300000"))
    ks = np.where(np.arange(nsk) % 10 == 0, np.arange(nsk) + rank * nsk, 7).astype(np.int32)
    tsk = HostRelation("sk", ["k", "v"], [P.INT4, P.INT8], [ks, np.arange(nsk, dtype=np.int64) + rank * nsk])
    dsk = capi.DeviceRelation.from_host(ctx, tsk)
    exs = capi.Executor(ctx, [dsk], motion=motion)
    scs = P.SeqScan(1, [("k", P.Var(1, 1, P.INT4)), ("v", P.Var(1, 2, P.INT8))])
    mhs = P.Motion(scs, P.MOTIONTYPE_HASH, [P.out_var(scs, 1)], world)
    aggs = P.Agg(mhs, P.AGG_PLAIN, P.AGGSPLIT_SIMPLE, [], [("n", P.Aggref(P.AGG_COUNT_STAR)), ("s", P.Aggref(P.AGG_SUM, P.out_var(mhs, 2)))])
    rsk = exs.run(P.Motion(aggs, P.MOTIONTYPE_GATHER))
    if rank == 0:
        tot_n = sum(int(r[0]) for r in rsk.rows)
        tot_s = sum(int(r[1]) for r in rsk.rows if r[1] is not None)
        big = max(int(r[0]) for r in rsk.rows)
        if tot_n != nsk * world or tot_s != sum(range(nsk * world)) or big < 0.9 * nsk * world:
            ok = False
            print("MULTIRANK FAIL: skewed Motion delivered %d rows (sum %d), largest segment %d" % (tot_n, tot_s, big))
        else:
            print("skewed Motion ok: %d rows, %d on one segment, repartitions %s" % (
                tot_n, big, [v.get("motion_repartitions") for v in rsk.instrument.values() if v.get("motion_repartitions")]))
    exs.close()
    dsk.free()
    # a segment that fails in the middle of a query: the others must get CBGPU_ERR_PEER, not wait for ever, and the
    # interconnect must be usable for the next query
    bad = HostRelation("bad", ["k", "v"], [P.INT4, P.INT8], [np.arange(64, dtype=np.int32), np.full(64, (1 << 62) if rank == world - 1 else 1, dtype=np.int64)])
    dbad = capi.DeviceRelation.from_host(ctx, bad)
    exb = capi.Executor(ctx, [dbad], motion=motion)
    scb = P.SeqScan(1, [("k", P.Var(1, 1, P.INT4)), ("v", P.Var(1, 2, P.INT8))])
    # v * 4 overflows int8 on the last segment only: its sender slice fails (CBGPU_ERR_OVERFLOW)
    mul = P.SeqScan(1, [("k", P.Var(1, 1, P.INT4)), ("v4", P.OpExpr(P.OP_MUL, P.Var(1, 2, P.INT8), P.Const(P.INT8, 4)))])
    mb = P.Motion(mul, P.MOTIONTYPE_HASH, [P.out_var(mul, 1)], world)
    code = None
    try:
        exb.run(P.Motion(mb, P.MOTIONTYPE_GATHER))
    except capi.CbgpuError as e:
        code = e.code
    want_code = -4 if rank == world - 1 else -7       # CBGPU_ERR_OVERFLOW / CBGPU_ERR_PEER
    if code != want_code:
        ok = False
        print("MULTIRANK FAIL: rank %d: failing segment gave code %r, expected %r" % (rank, code, want_code))
    # ... and the next query works
    rb = exb.run(P.Motion(P.Motion(scb, P.MOTIONTYPE_HASH, [P.out_var(scb, 1)], world), P.MOTIONTYPE_GATHER))
    if rank == 0 and len(rb.rows) != 64 * world:
        ok = False
        print("MULTIRANK FAIL: query after a failed one returned %d rows" % len(rb.rows))
    exb.close()
    dbad.free()
    # sorted Gather (Motion.sendSorted): every segment sends its local top 30 by (count desc, k), the receiver merges the
    # streams on the device - between processes the rows arrive through the window's per-sender slots (or NCCL)
    nsg = 20011
    rng = np.random.default_rng(77 + rank)
    tsg = HostRelation("sg", ["k", "v"], [P.INT4, P.INT8], [rng.integers(0, 300, nsg).astype(np.int32), rng.integers(0, 1000, nsg)])
    dsg = capi.DeviceRelation.from_host(ctx, tsg)
    exg = capi.Executor(ctx, [dsg], motion=motion)
    scg = P.SeqScan(1, [("k", P.Var(1, 1, P.INT4)), ("v", P.Var(1, 2, P.INT8))])
    agg_g = P.Agg(scg, P.AGG_HASHED, P.AGGSPLIT_SIMPLE, [1], [("k", P.out_var(scg, 1)), ("s", P.Aggref(P.AGG_SUM, P.out_var(scg, 2))),
                                                                 ("n", P.Aggref(P.AGG_COUNT_STAR))], num_groups=512)
    keys_g = [(3, True), (1, False)]
    rsg = exg.run(P.Motion(P.LimitSort(agg_g, keys_g, 30), P.MOTIONTYPE_GATHER, sort_keys=keys_g))
    if rank == 0:
        order = [(-int(r[2]), int(r[0])) for r in rsg.rows]
        if len(rsg.rows) != 30 * world or order != sorted(order):
            ok = False
            print("MULTIRANK FAIL: sorted Gather returned %d rows, in order: %s" % (len(rsg.rows), order == sorted(order)))
    exg.close()
    dsg.free()
    r1 = ex.run(tpch.q1_plan(world))
    r3 = ex.run(tpch.q3_plan(seg, world, customer_replicated=replicated, merge_gather=True))
    r5 = ex.run(tpch.q5_plan(reg, world, replicated=replicated))
    if rank == 0:
        ok = ok and tpch.format_q1(r1.rows) == exp["q1"]
        ok = ok and tpch.format_q3(r3.rows) == exp["q3"]
        ok = ok and tpch.format_q5(r5.rows, exp["dict"]["n_name_dict"]) == exp["q5"]
        print("MULTIRANK", "PASS" if ok else "FAIL", "world", world, "replicated", replicated, "direct" if motion.direct() else "staged",
              "motion bytes sent by rank 0", motion.bytes_sent(), "host syncs", motion.host_syncs(), "collectives", motion.collectives())
        if not ok:
            print(tpch.format_q1(r1.rows), tpch.format_q3(r3.rows), tpch.format_q5(r5.rows, exp["dict"]["n_name_dict"]))
    else:
        ok = ok and len(r1.rows) == 0 and len(r3.rows) == 0 and len(r5.rows) == 0     # only the gather receiver emits
        if not ok:
            print("MULTIRANK FAIL: non-root rank emitted rows")
    flag = torch.tensor([0 if ok else 1], device="cpu" if boot_gloo else "cuda")
    dist.all_reduce(flag)
    ex.close()
    motion.close()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    main()
