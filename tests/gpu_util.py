"""Helpers shared by the GPU parity tests."""
import ctypes as C

import numpy as np

from cloudberry_b200 import capi, tpch
from cloudberry_b200 import plan as P


def shard(oracle, rels, nsegs, dist=None):
    """Per-segment range tables, distributed as the reference does (cdbhash of the distribution key,
    jump consistent hash; undistributed relations replicated).  Placement computed by the oracle."""
    dist = tpch.DIST_KEY if dist is None else dist
    L = oracle.lib()
    segs = [[] for _ in range(nsegs)]
    for rel in rels:
        key = dist.get(rel.name)
        if key is None:
            for s in range(nsegs):
                segs[s].append(rel)
            continue
        a = rel.attno(key) - 1
        t = (C.c_int32 * 1)(rel.types[a])
        dest = np.array([L.ora_cdbhash_segment(t, (C.c_int64 * 1)(int(v)), None, 1, nsegs) for v in rel.columns[a]])
        for s in range(nsegs):
            segs[s].append(rel.take(np.nonzero(dest == s)[0]))
    return segs


def to_device(ctx, rels):
    return [capi.DeviceRelation.from_host(ctx, r) for r in rels]


def canon(rows):
    """Order-insensitive comparison form (hash aggregation emits in table order, nodeAgg.c:3053)."""
    return sorted([tuple("NULL" if v is None else v for v in r) for r in rows], key=lambda r: tuple(map(str, r)))
