"""character(n) / varchar columns decoded ON THE DEVICE into dictionary codes (cbgpu_dict: cbgpu_aocs_dict_collect ->
cbgpu_dict_finalize -> cbgpu_aocs_decode_dict_column) == the strings the reference's block writer was given
(tests/golden/aocs_text_columns.npz), then joins / groupings over such columns against the oracle."""
import numpy as np
import pytest

from cloudberry_b200 import capi
from cloudberry_b200 import plan as P
from test_aocs_format import TCASES

pytestmark = pytest.mark.gpu
COMP = {"": 0, "zlib": 1, "zstd": 2}


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def canon_text(v, bpchar):
    return v.rstrip(b" ") if bpchar else v


@pytest.mark.parametrize("case", TCASES, ids=[c[0] for c in TCASES])
def test_codes_spell_the_values_that_went_in(ctx, case):
    name, typname, checksum, blocksize, nblocks, raw, values, nulls, comp = case
    bpchar = typname == "bpchars"
    d = capi.DeviceDict(ctx, max_entries=8192, bpchar=bpchar)
    d.collect(raw, checksum, COMP[comp])
    texts = d.entries()
    want = sorted(set(canon_text(v, bpchar) for v, z in zip(values, nulls) if not z))
    assert texts == want                                   # distinct values, in byte order: code order == string order
    rel = capi.DeviceRelation(ctx, len(values) + 3, [P.DICT32])
    assert rel.load_aocs_dict_column(0, raw, checksum, d, COMP[comp], row_offset=3) == len(values)
    got, gotnull = rel.read_column(0, 3, 3 + len(values))
    assert np.array_equal(gotnull.astype(np.uint8), nulls)
    keep = nulls == 0
    assert [texts[c] for c in got[keep]] == [canon_text(v, bpchar) for v, z in zip(values, nulls) if not z]
    # host-side lookups for plan constants: blanks at the end do not count for character(n), do for varchar
    for t in want[:5]:
        assert d.lookup(t) == texts.index(t)
        assert d.lookup(t + b"  ") == (texts.index(t) if bpchar else (texts.index(t + b"  ") if t + b"  " in texts else -1))
    assert d.lookup(b"no such value") == -1
    if len(want) <= 256:
        r8 = capi.DeviceRelation(ctx, len(values), [P.DICT8])
        assert r8.load_aocs_dict_column(0, raw, checksum, d, COMP[comp]) == len(values)
        g8, _ = r8.read_column(0)
        assert np.array_equal(g8[keep].astype(np.int64), got[keep].astype(np.int64))
        r8.free()
    else:
        r8 = capi.DeviceRelation(ctx, len(values), [P.DICT8])
        with pytest.raises(capi.CbgpuError) as e:
            r8.load_aocs_dict_column(0, raw, checksum, d, COMP[comp])
        assert e.value.code == -4
        r8.free()
    rel.free()
    d.free()


def test_one_dictionary_over_several_files_and_limits(ctx):
    by = {c[0]: c for c in TCASES}
    a, b = by["shipmode_bpchar10_nulls"], by["shipmode_rle"]
    d = capi.DeviceDict(ctx, max_entries=64)
    d.collect(a[5], a[2])
    d.collect(b[5], b[2])
    d.collect(a[5], a[2])                                  # seeing a file twice adds nothing
    assert d.entries() == sorted(m.encode() for m in ["REG AIR", "AIR", "RAIL", "SHIP", "TRUCK", "MAIL", "FOB"])
    with pytest.raises(capi.CbgpuError):
        d.collect(a[5], a[2])                              # finalized
    # a value the dictionary never saw: reported, not given some code
    other = by["mktsegment_zlib5"]
    rel = capi.DeviceRelation(ctx, len(other[6]), [P.DICT8])
    with pytest.raises(capi.CbgpuError) as e:
        rel.load_aocs_dict_column(0, other[5], other[2], d, 1)
    assert e.value.code == -2
    rel.free()
    d.free()
    # more distinct values than the dictionary may hold
    many = by["varchar_many_distinct_zstd"]
    small = capi.DeviceDict(ctx, max_entries=100, bpchar=False)
    with pytest.raises(capi.CbgpuError) as e:
        small.collect(many[5], many[2], 2)
        small.finalize()
    assert e.value.code == -5
    small.free()


def test_group_and_join_on_dictionary_columns(ctx, oracle):
    """lineitem-like rows grouped by l_shipmode, joined to a small table on the mode (a DICT8 hash key: per-code
    hashbpchar from the dictionary), both sides decoded from column files"""
    from test_aocs_format import CASES
    from cloudberry_b200.relation import HostRelation
    from cloudberry_b200.tpch import _child_var
    from gpu_util import canon
    by = {c[0]: c for c in TCASES}
    ship = by["shipmode_bpchar10_nulls"]
    price = {c[0]: c for c in CASES}["numeric_price"]
    n = len(ship[6])
    assert len(price[7]) == n
    d = capi.DeviceDict(ctx, max_entries=64)
    d.collect(ship[5], ship[2])
    texts = d.entries()
    rel = capi.DeviceRelation(ctx, n, [P.DICT8, P.NUMERIC])
    assert rel.load_aocs_dict_column(0, ship[5], ship[2], d) == n
    assert rel.load_aocs_column(1, price[6], price[2], -1, 1, 4) == n
    codes = np.array([0 if z else texts.index(v.rstrip(b" ")) for v, z in zip(ship[6], ship[7])], dtype=np.uint8)
    host = HostRelation("l", ["m", "p"], [P.DICT8, P.NUMERIC], [codes, price[7]], nulls=[ship[7], None],
                        dict_texts=[[t.decode() for t in texts], None])
    host.set_dict_hashes(oracle.hashbpchar)
    sc = P.SeqScan(1, [("m", P.Var(1, 1, P.DICT8)), ("p", P.Var(1, 2, P.NUMERIC, 2))])
    v = _child_var(sc)
    plan = P.Agg(sc, P.AGG_HASHED, P.AGGSPLIT_SIMPLE, [1], [("m", v("m")), ("s", P.Aggref(P.AGG_SUM, v("p"))), ("n", P.Aggref(P.AGG_COUNT_STAR))],
                 num_groups=16)
    ex = capi.Executor(ctx, [rel])
    got = ex.run(plan)
    want = oracle.execute(plan, [[host]])
    assert canon(got.rows) == canon(want.rows) and len(want.rows) == 8          # seven modes + the NULL group
    ex.close()
    rel.free()
    d.free()
