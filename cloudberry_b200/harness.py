"""Multi-process harness pieces shared by bench.py, the multi-GPU worker and the CPU (gloo) tests.

One process per GPU-segment, launched by torchrun.  torch.distributed is plumbing only (rendezvous,
barrier, max-over-ranks of timings, handing the NCCL rendezvous token around); the data path's
collectives live in csrc/motion.cu.  `exchange_by_destination` is the reference implementation of
the Redistribute protocol (count exchange, then payload all-to-all) on host tensors: the CPU tests
run it over gloo to pin the protocol the CUDA side follows."""
import os

import numpy as np


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_dist(backend, local_rank=0):
    import torch
    import torch.distributed as dist
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group(backend)
    return dist


def broadcast_token(dist, make_token, rank):
    """Rank 0 makes the interconnect's rendezvous token (cbgpu_motion_unique_id), everyone gets it."""
    box = [make_token() if rank == 0 else None]
    if dist is not None:
        dist.broadcast_object_list(box, src=0)
    return box[0]


def shard_rows(rank, nrows_per_rank):
    """Weak-scaling shard of a counter-generated table: rank r owns rows [r * n, (r + 1) * n)."""
    return rank * nrows_per_rank, (rank + 1) * nrows_per_rank


def max_over_ranks(dist, values, device="cpu"):
    import torch
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def exchange_by_destination(dist, rank, world, dest, columns):
    """Redistribute rows to their destination ranks.
    dest: int array (nrows) of destination ranks; columns: list of numpy arrays (nrows each).
    Protocol = csrc/motion.cu: (1) every rank learns every rank's per-destination counts (all-gather),
    (2) one all-to-all per column of exactly those row ranges, sender-major on the receiver.
    Returns the received columns."""
    import torch
    order = np.argsort(dest, kind="stable")
    counts = np.bincount(dest, minlength=world).astype(np.int64)
    mine = torch.from_numpy(counts.copy())
    gathered = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, mine)
    matrix = torch.stack(gathered).numpy()            # matrix[s][d]
    recv_counts = matrix[:, rank]
    out = []
    for col in columns:
        col = np.ascontiguousarray(col[order])
        send = torch.from_numpy(col.view(np.uint8).reshape(len(col), -1).copy()) if len(col) else torch.zeros((0, col.dtype.itemsize), dtype=torch.uint8)
        w = col.dtype.itemsize
        send_list = list(torch.split(send.reshape(-1), [int(c) * w for c in counts]))
        recv_list = [torch.zeros(int(c) * w, dtype=torch.uint8) for c in recv_counts]
        dist.all_to_all(recv_list, send_list) if dist.get_backend() != "gloo" else _gloo_all_to_all(dist, rank, world, recv_list, send_list)
        out.append(torch.cat(recv_list).numpy().view(col.dtype) if recv_counts.sum() else np.zeros(0, dtype=col.dtype))
    return out, matrix


def _gloo_all_to_all(dist, rank, world, recv_list, send_list):
    """gloo has no all_to_all for uneven splits on every build: pairwise isend/irecv."""
    reqs = []
    for peer in range(world):
        if peer == rank:
            recv_list[peer].copy_(send_list[peer])
            continue
        if send_list[peer].numel():
            reqs.append(dist.isend(send_list[peer].contiguous(), dst=peer))
        if recv_list[peer].numel():
            reqs.append(dist.irecv(recv_list[peer], src=peer))
    for r in reqs:
        r.wait()


def device_tables(ctx, sf, seed=42, lineitem=None):
    """All six TPC-H relations of the queries on this path, generated on the device (csrc/gen.cu: the same
    counter-based formulas as tpch.py) in range-table order; `lineitem` reuses an existing relation."""
    from . import capi, tpch
    sz = tpch.sizes(int(sf) if float(sf).is_integer() else sf)
    G = ctx.L
    rels = {}
    for name in tpch.RT:
        if name == "lineitem" and lineitem is not None:
            rels[name] = lineitem
            continue
        types = [t for _, t in tpch.SCHEMA[name]]
        rels[name] = capi.DeviceRelation(ctx, sz[name], types, name=name)
    if lineitem is None:
        ctx.check(G.cbgpu_gen_lineitem(ctx.h, rels["lineitem"].h, seed, 0, sz["supplier"], sz["part"]))
    ctx.check(G.cbgpu_gen_orders(ctx.h, rels["orders"].h, seed, 0, sz["customer"]))
    ctx.check(G.cbgpu_gen_customer(ctx.h, rels["customer"].h, seed))
    ctx.check(G.cbgpu_gen_supplier(ctx.h, rels["supplier"].h, seed))
    nation, region = tpch.gen_nation_region()
    rels["nation"].load(tpch._rel("nation", nation, {"n_name": tpch.NATIONS}).set_dict_hashes(capi.hashbpchar))
    rels["region"].load(tpch._rel("region", region, {"r_name": tpch.REGIONS}).set_dict_hashes(capi.hashbpchar))
    rels["customer"].set_dict_hash(2, np.array([capi.hashbpchar(s) for s in tpch.SEGMENTS], dtype=np.uint32))
    ctx.sync()
    return [rels[n] for n in tpch.RT], sz


# projected bytes per row of every base table a query scans (each projected column once)
QUERY_BYTES = {
    "q3": {"lineitem": 28, "orders": 20, "customer": 5},
    "q5": {"lineitem": 28, "orders": 16, "customer": 8, "supplier": 8, "nation": 9, "region": 5},
}


def query_rows_bytes(q, sz):
    rows = sum(sz[t] for t in QUERY_BYTES[q])
    nbytes = sum(sz[t] * w for t, w in QUERY_BYTES[q].items())
    return rows, nbytes


def distribute_by_hash(ctx, motion, rel, keycol, name):
    """`DISTRIBUTED BY (key)` at load time: every rank holds a row-range slice `rel` of the table; a
    Redistribute Motion (cdbhash + jump consistent hash of the key, NCCL all-to-all) leaves each rank with
    the rows the reference would store on that segment.  Returns the shard as a device relation."""
    from . import capi
    from . import plan as P
    ex = capi.Executor(ctx, [rel], motion=motion)
    tl = [("c%d" % i, P.Var(1, i + 1, t, rel.dscales[i])) for i, t in enumerate(rel.types)]
    scan = P.SeqScan(1, tl)
    m = P.Motion(scan, P.MOTIONTYPE_HASH, [P.out_var(scan, keycol + 1)], motion.nranks)
    out = ex.run_batch(m, name)
    ex.close()
    return out


def distributed_tables(ctx, motion, sf, rank, world, seed=42, max_piece_rows=150_000_000):
    """The six relations of an SF-sized database spread over `world` GPU-segments the way the reference's
    DDL would: lineitem and orders by orderkey, customer by c_custkey, supplier by s_suppkey, nation and
    region replicated.  Each rank generates a 1/world row range of every distributed table, then the
    load-time Redistribute moves the rows to their segments."""
    from . import capi, tpch
    sz = tpch.sizes(int(sf) if float(sf).is_integer() else sf)
    G = ctx.L
    keys = {"lineitem": 0, "orders": 0, "customer": 0, "supplier": 0}
    shards = {}
    seg_hash = np.array([capi.hashbpchar(s) for s in tpch.SEGMENTS], dtype=np.uint32)
    for name in ("lineitem", "orders", "customer", "supplier"):
        n = sz[name]
        lo, hi = rank * n // world, (rank + 1) * n // world
        types = [t for _, t in tpch.SCHEMA[name]]
        # load in pieces: a piece's slice, its Motion buffers and its share of the shard are all that is in flight, so a
        # table far larger than the peer-memory window (SF300 lineitem on 2 GPUs: 50 GB per rank) loads without ever holding
        # slice + send buffer + receive buffer of the whole table at once.  Every rank cuts the same number of pieces.
        piece_rows = max_piece_rows
        npieces = max(1, -(-max(((r + 1) * n // world - r * n // world) for r in range(world)) // piece_rows))
        pieces = []
        for k in range(npieces):
            plo, phi = lo + (hi - lo) * k // npieces, lo + (hi - lo) * (k + 1) // npieces
            sl = capi.DeviceRelation(ctx, phi - plo, types, name=name + "_slice")
            if name == "lineitem":
                ctx.check(G.cbgpu_gen_lineitem(ctx.h, sl.h, seed, plo, sz["supplier"], sz["part"]))
            elif name == "orders":
                ctx.check(G.cbgpu_gen_orders(ctx.h, sl.h, seed, plo, sz["customer"]))
            elif name == "customer":
                ctx.check(G.cbgpu_gen_customer_range(ctx.h, sl.h, seed, plo))
                sl.set_dict_hash(2, seg_hash)
            else:
                ctx.check(G.cbgpu_gen_supplier_range(ctx.h, sl.h, seed, plo))
            ctx.sync()
            pieces.append(distribute_by_hash(ctx, motion, sl, keys[name], name))
            sl.free()
        if len(pieces) == 1:
            shards[name] = pieces[0]
        else:
            total = sum(p.nrows for p in pieces)
            whole = capi.DeviceRelation(ctx, total, types, name=name)
            at = 0
            for p in pieces:
                if p.nrows:
                    ctx.check(G.cbgpu_rel_copy_rows(whole.h, at, p.h, 0, p.nrows))
                at += p.nrows
                p.free()
            if name == "customer":
                whole.set_dict_hash(2, seg_hash)
            shards[name] = whole
    nation, region = tpch.gen_nation_region()
    for name, cols, texts in (("nation", nation, {"n_name": tpch.NATIONS}), ("region", region, {"r_name": tpch.REGIONS})):
        types = [t for _, t in tpch.SCHEMA[name]]
        r = capi.DeviceRelation(ctx, sz[name], types, name=name)
        r.load(tpch._rel(name, cols, texts).set_dict_hashes(capi.hashbpchar))
        shards[name] = r
    ctx.sync()
    return [shards[n] for n in tpch.RT], sz
