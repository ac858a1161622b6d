"""The checker checked (not gpu): the oracle's SQL semantics on the edge cases the GPU parity tests lean on it for
(tests/test_gpu_edge.py) - NULL join keys never match (strict hash operators, nodeHash.c:2161), NULL group keys form one
group (execGrouping.c:548), invisible rows do not exist, duplicate build keys multiply matches, LEFT / SEMI / ANTI joins,
count / sum / avg / min / max over NULLs - against a from-first-principles evaluation in plain Python (nested dictionaries,
no shared code with either executor).  avg(bigint)'s numeric text comes from the reference's numeric_div where oracle/_ref
is built."""
import ctypes as C
import os

import numpy as np
import pytest

from cloudberry_b200 import plan as P
from test_gpu_edge import agg_over, dim, fact, scan

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "..", "oracle", "_ref", "libexec_ref.so")


def _dec(v, ds):
    s = "-" if v < 0 else ""
    d = str(abs(v)).rjust(ds + 1, "0")
    return s + (d[:-ds] + "." + d[-ds:] if ds else d)


def _avg_text(total, n):
    if not os.path.exists(SO):
        return None
    R = C.CDLL(SO)
    R.ref_numeric_binop.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
    buf = C.create_string_buffer(256)
    assert R.ref_numeric_binop(3, str(total).encode(), str(n).encode(), buf, 256) == 0
    return buf.value.decode()


def _rows(rel):
    """visible rows as dicts, None for NULL"""
    n = rel.nrows
    vis = np.ones(n, dtype=bool) if rel.visimap is None else np.unpackbits(rel.visimap, bitorder="little")[:n].astype(bool)
    out = []
    for i in range(n):
        if not vis[i]:
            continue
        out.append({nm: (None if rel.nulls[c] is not None and rel.nulls[c][i] else rel.columns[c][i].item())
                    for c, nm in enumerate(rel.names)})
    return out


def _expected(fo, do, jointype):
    build = {}
    for d in _rows(do):
        if d["dk"] is not None:                 # a NULL build key is dropped (nodeHash.c:2161)
            build.setdefault(d["dk"], []).append(d)
    joined = []
    for f in _rows(fo):
        matches = build.get(f["k"], []) if f["k"] is not None else []
        if jointype == P.JOIN_INNER:
            joined += [dict(f, w=m["w"], c=m["c"]) for m in matches]
        elif jointype == P.JOIN_LEFT:
            joined += [dict(f, w=m["w"], c=m["c"]) for m in matches] or [dict(f, w=None, c=None)]
        elif jointype == P.JOIN_SEMI:
            joined += [f] if matches else []
        else:
            joined += [] if matches else [f]
    with_inner = jointype in (P.JOIN_INNER, P.JOIN_LEFT)
    groups = {}
    for r in joined:
        key = (r["g"], r["c"]) if with_inner else (r["g"],)
        groups.setdefault(key, []).append(r)
    out = []
    for key, rs in groups.items():
        vs = [r["v"] for r in rs if r["v"] is not None]
        row = list(key) + [_dec(sum(r["amt"] for r in rs), 2), len(rs), len(vs),
                           _avg_text(sum(vs), len(vs)) if vs else None, min(vs) if vs else None]
        if with_inner:
            ws = [r["w"] for r in rs if r["w"] is not None]
            row += [max(ws) if ws else None, str(sum(ws)) if ws else None]
        out.append(row)
    return out


def _canon(rows):
    return sorted([tuple("NULL" if v is None else v for v in r) for r in rows], key=lambda r: tuple(map(str, r)))


@pytest.mark.parametrize("jointype", [P.JOIN_INNER, P.JOIN_LEFT, P.JOIN_SEMI, P.JOIN_ANTI])
@pytest.mark.parametrize("nf,nd,dup,null_frac,visible", [(3000, 40, 1, 0.0, 1.0), (3000, 90, 3, 0.1, 0.8), (0, 40, 1, 0.0, 1.0),
                                                          (500, 0, 1, 0.1, 1.0), (2049, 1, 1, 0.3, 0.5)])
def test_join_and_aggregate_semantics(oracle, jointype, nf, nd, dup, null_frac, visible):
    fo = fact(nf, seed=5, null_frac=null_frac, visible_frac=visible).set_dict_hashes(oracle.hashbpchar)
    do = dim(nd, seed=6, null_frac=null_frac, dup=dup).set_dict_hashes(oracle.hashbpchar)
    sf = scan(1, fo, ["k", "v", "amt", "g"])
    sd = scan(2, do, ["dk", "w", "c"])
    h = P.Hash(sd, [P.out_var(sd, 1)])
    targets = [("g", P.out_var(sf, 4)), ("v", P.out_var(sf, 2)), ("amt", P.out_var(sf, 3))]
    if jointype in (P.JOIN_INNER, P.JOIN_LEFT):
        targets += [("w", P.InnerVar(2, P.INT8)), ("c", P.InnerVar(3, P.DICT8))]
    j = P.HashJoin(jointype, sf, h, [P.out_var(sf, 1)], targets)
    names = [t[0] for t in targets]
    aggs = [("s", P.AGG_SUM, "amt"), ("n", P.AGG_COUNT_STAR, None), ("cv", P.AGG_COUNT, "v"), ("av", P.AGG_AVG, "v"),
            ("mn", P.AGG_MIN, "v")]
    if "w" in names:
        aggs += [("mx", P.AGG_MAX, "w"), ("sw", P.AGG_SUM, "w")]
    plan = agg_over(j, names, ["g"] + (["c"] if "c" in names else []), aggs)
    got = oracle.execute(plan, [[fo, do]]).rows
    want = _expected(fo, do, jointype)
    if not os.path.exists(SO):                   # no reference numeric_div here: leave the avg column out
        av = len(want[0]) - (3 if "w" in names else 1) - 1 if want else 0
        got = [r[:av] + r[av + 1:] for r in got]
        want = [r[:av] + r[av + 1:] for r in want]
    assert _canon(got) == _canon(want)


def test_float8_sum_is_sequential_float8pl(oracle):
    """sum(float8) = float8pl row after row (float.c:774), avg = sum / N (float8_avg, float.c:3148): the oracle's doubles
    are the ones a sequential executor produces, bit for bit - the GPU's tree-ordered sums are held to 1e-12 sqrt(N) of these"""
    n = 50021
    fo = fact(n, seed=7).set_dict_hashes(oracle.hashbpchar)
    sc = scan(1, fo, ["g", "x"])
    plan = agg_over(sc, ["g", "x"], ["g"], [("sx", P.AGG_SUM, "x"), ("ax", P.AGG_AVG, "x"), ("n", P.AGG_COUNT_STAR, None)])
    got = {r[0]: r for r in oracle.execute(plan, [[fo]]).rows}
    acc, cnt = {}, {}
    g, x = fo.columns[fo.names.index("g")], fo.columns[fo.names.index("x")]
    for i in range(n):
        k = int(g[i])
        acc[k] = acc.get(k, 0.0) + float(x[i])
        cnt[k] = cnt.get(k, 0) + 1
    assert set(got) == set(acc)
    for k in acc:
        assert float(got[k][1]) == acc[k]
        assert float(got[k][2]) == acc[k] / cnt[k]
        assert got[k][3] == cnt[k]
