"""The oracle's RIGHT / FULL hash joins (oracle/oracle.c hashjoin_next: HJ_FILL_INNER_TUPLES, nodeHashjoin.c:676-706;
keep_nulls, nodeHash.c:2171-2190) against a brute-force restatement of SQL's outer join in plain Python - the check that
the checker is right before the CUDA path is compared with it (tests/test_gpu_capacity.py)."""
import pytest

from cloudberry_b200 import plan as P
from test_gpu_capacity import _outer_join_plan
from test_gpu_edge import dim, fact


def _brute(jointype, fo, do):
    def col(rel, name):
        i = rel.names.index(name)
        vals = rel.columns[i].tolist()
        nl = rel.nulls[i].tolist() if rel.nulls[i] is not None else [0] * len(vals)
        return [None if n else v for v, n in zip(vals, nl)]
    k, g, amt = col(fo, "k"), col(fo, "g"), col(fo, "amt")
    dk, w, c = col(do, "dk"), col(do, "w"), col(do, "c")
    rows, seen = [], set()
    for i in range(len(k)):
        hit = False
        for j in range(len(dk)):
            if k[i] is not None and dk[j] is not None and k[i] == dk[j]:
                rows.append((k[i], g[i], amt[i], dk[j], w[j], c[j]))
                seen.add(j)
                hit = True
        if not hit and jointype in (P.JOIN_LEFT, P.JOIN_FULL):
            rows.append((k[i], g[i], amt[i], None, None, None))
    if jointype in (P.JOIN_RIGHT, P.JOIN_FULL):
        rows += [(None, None, None, dk[j], w[j], c[j]) for j in range(len(dk)) if j not in seen]
    return rows


@pytest.mark.parametrize("jointype", [P.JOIN_LEFT, P.JOIN_RIGHT, P.JOIN_FULL])
@pytest.mark.parametrize("nf,nd,dup", [(0, 20, 1), (50, 0, 1), (200, 40, 1), (300, 60, 3)])
def test_outer_joins_against_brute_force(jointype, nf, nd, dup):
    from oracle import oracle as O
    fo = fact(nf, seed=41, null_frac=0.15, kmax=60).set_dict_hashes(O.hashbpchar)
    do = dim(nd, seed=42, null_frac=0.15, dup=dup, kmax=60).set_dict_hashes(O.hashbpchar)
    got = O.execute(_outer_join_plan(jointype, fo, do), [[fo, do]]).rows
    want = _brute(jointype, fo, do)

    from decimal import Decimal

    def norm(rows):
        return sorted(tuple("~" if x is None else str(x) for x in r) for r in rows)
    # the oracle returns numerics as decimal strings at the column's scale (2) and dictionary columns as their codes
    want = [(a, b, None if cc is None else str(Decimal(cc).scaleb(-2)), d, e, f) for a, b, cc, d, e, f in want]
    assert norm(got) == norm(want)
    assert len(want) > 0 or nd == 0 or nf == 0


def _notin_plan(fo, do):
    from test_gpu_edge import scan
    sf = scan(1, fo, ["k", "amt", "g"])
    sd = scan(2, do, ["dk"])
    h = P.Hash(sd, [P.out_var(sd, 1)])
    return P.HashJoin(P.JOIN_LASJ_NOTIN, sf, h, [P.out_var(sf, 1)], [("k", P.out_var(sf, 1)), ("amt", P.out_var(sf, 2)), ("g", P.out_var(sf, 3))])


@pytest.mark.parametrize("nf,nd,fnull,dnull", [(300, 40, 0.15, 0.0), (300, 40, 0.15, 0.2), (300, 0, 0.15, 0.0), (0, 40, 0.0, 0.0), (300, 40, 0.0, 0.0)])
def test_not_in_join_against_sql_semantics(nf, nd, fnull, dnull):
    """k NOT IN (select dk ...): LASJ_NOTIN (nodeHashjoin.c:371-390, 578-590).  A NULL in the set makes the predicate unknown
    for every row; a NULL k is unknown unless the set is empty; otherwise an anti join."""
    from oracle import oracle as O
    from decimal import Decimal
    fo = fact(nf, seed=43, null_frac=fnull, kmax=60).set_dict_hashes(O.hashbpchar)
    do = dim(nd, seed=44, null_frac=dnull, kmax=60).set_dict_hashes(O.hashbpchar)

    def col(rel, name):
        i = rel.names.index(name)
        vals = rel.columns[i].tolist()
        nl = rel.nulls[i].tolist() if rel.nulls[i] is not None else [0] * len(vals)
        return [None if n else v for v, n in zip(vals, nl)]
    k, amt, g = col(fo, "k"), col(fo, "amt"), col(fo, "g")
    dk = col(do, "dk")
    if any(x is None for x in dk):
        want = []
    elif not dk:
        want = list(zip(k, amt, g))
    else:
        keys = set(dk)
        want = [(a, b, c) for a, b, c in zip(k, amt, g) if a is not None and a not in keys]
    got = O.execute(_notin_plan(fo, do), [[fo, do]]).rows
    norm = lambda rows: sorted(tuple("~" if x is None else str(x) for x in r) for r in rows)
    assert norm(got) == norm([(a, None if b is None else str(Decimal(b).scaleb(-2)), c) for a, b, c in want])
    if dnull == 0 and nd and nf:
        assert 0 < len(want) < nf
