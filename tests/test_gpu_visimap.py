"""pg_aovisimap entries applied ON THE DEVICE (cbgpu_aocs_apply_visimap: k_visimap_expand / k_visimap_apply) == the rows
that were hidden when the REFERENCE's Bitmap_Compress wrote the entries (tests/golden/aocs_visimap.npz), then a query over
a relation whose visibility came that way."""
import numpy as np
import pytest

from cloudberry_b200 import capi
from cloudberry_b200 import plan as P
from test_visimap_format import VCASES

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("case", VCASES, ids=[c[0] for c in VCASES])
@pytest.mark.parametrize("row_offset", [0, 13])
def test_device_visibility_matches_hidden_rows(ctx, case, row_offset):
    name, raw, checksum, entries, visible = case
    n = len(visible)
    rel = capi.DeviceRelation(ctx, n + row_offset + 3, [P.INT4])
    rng = np.random.default_rng(3)
    shuffled = [entries[i] for i in rng.permutation(len(entries))]            # any order
    hidden = rel.apply_visimap(raw, checksum, shuffled, row_offset=row_offset)
    assert hidden == int((~visible).sum())
    got = rel.read_visimap()
    assert np.array_equal(got[row_offset:row_offset + n], visible)
    assert got[:row_offset].all() and got[row_offset + n:].all()               # rows outside the file keep their state
    # applying again (e.g. after a VACUUM rewrote the entries) replaces the state of the range
    assert rel.apply_visimap(raw, checksum, [], row_offset=row_offset) == 0
    assert rel.read_visimap().all()
    rel.free()


def test_malformed_entries_are_reported(ctx):
    name, raw, checksum, entries, visible = VCASES[0]
    rel = capi.DeviceRelation(ctx, len(visible), [P.INT4])
    first, good = entries[0]
    for bad in (b"\x02\0\0\0" + good[4:],                        # version
                good[:len(good) // 2],                           # bit stream ends early
                b"\x01\0\0\0" + bytes([0x80, 0x01, 0x80]),       # repeat token before any block
                good[:4] + bytes([good[4] | 0x0F, 0xFF]) + good[6:]):      # block count 4095
        with pytest.raises(capi.CbgpuError) as e:
            rel.apply_visimap(raw, checksum, [(first, bad)] + entries[1:])
        assert e.value.code == -6
    with pytest.raises(capi.CbgpuError):
        rel.apply_visimap(raw, checksum, [(first + 5, good)])                  # not a multiple of 32768
    with pytest.raises(capi.CbgpuError):
        rel.apply_visimap(raw, checksum, [(first, good), (first, good)])       # the same range twice
    assert rel.apply_visimap(raw, checksum, entries) == int((~visible).sum())
    rel.free()


def test_query_sees_only_visible_rows(ctx, oracle):
    """decode a column file, apply its visimap entries, count / sum: the oracle gets the same bitmap"""
    from test_aocs_format import CASES
    from cloudberry_b200.relation import HostRelation
    from cloudberry_b200.tpch import _child_var
    from gpu_util import canon
    name, raw, checksum, entries, visible = {c[0]: c for c in VCASES}["long_run_scattered"]
    col = {c[0]: c for c in CASES}["rle_numeric_long_run"]
    values = col[7]
    rel = capi.DeviceRelation(ctx, len(values), [P.NUMERIC], dscales=[2])
    assert rel.load_aocs_column(0, raw, checksum, -1, 1, 4) == len(values)
    assert rel.apply_visimap(raw, checksum, entries) == int((~visible).sum())
    host = HostRelation("t", ["p"], [P.NUMERIC], [values], dscales=[2], visimap=np.packbits(visible, bitorder="little"))
    sc = P.SeqScan(1, [("p", P.Var(1, 1, P.NUMERIC, 2))])
    v = _child_var(sc)
    plan = P.Agg(sc, P.AGG_PLAIN, P.AGGSPLIT_SIMPLE, [], [("s", P.Aggref(P.AGG_SUM, v("p"))), ("n", P.Aggref(P.AGG_COUNT_STAR))])
    ex = capi.Executor(ctx, [rel])
    got = ex.run(plan)
    want = oracle.execute(plan, [[host]])
    assert canon(got.rows) == canon(want.rows)
    assert int(want.rows[0][1]) == int(visible.sum())
    ex.close()
    rel.free()


def test_segment_files_from_disk(ctx, oracle, tmp_path):
    """a two-column table's segment file 1 on disk under the reference's file names, bytes of an aborted append behind the
    recorded EOF, its pg_aovisimap rows: cb_aocs_load_segfile reads, decodes, hides; a query over it == the oracle"""
    from test_aocs_format import CASES
    from cloudberry_b200.relation import HostRelation
    from cloudberry_b200.tpch import _child_var
    from gpu_util import canon
    by = {c[0]: c for c in CASES}
    price, flags = by["numeric_price"], by["bpchar1_flags"]                      # the same 20011 rows, two columns
    n = len(price[7])
    base = str(tmp_path / "24576")
    p1, p2 = capi.aocs_segfile_path(base, 1, 1), capi.aocs_segfile_path(base, 1, 2)
    assert p1.endswith("24576.1") and p2.endswith("24576.129")
    open(p1, "wb").write(flags[6] + b"\xde\xad\xbe\xef" * 100)                  # garbage past the EOF must not be read
    open(p2, "wb").write(price[6])
    hidden_rows = np.unique(np.random.default_rng(8).integers(1, n + 1, 700))     # row numbers 1..n (first_row_no 0 covers them)
    from oracle import aocs_format as A
    if A.ref_lib() is not None:
        payload = A.ref_visimap_entry(hidden_rows)
    else:
        # no reference library on this box: the uncompressed entry type, assembled by hand (bitmap_compression.c:118-123)
        words = np.zeros(1024, dtype="<u4")
        np.bitwise_or.at(words, hidden_rows // 32, (1 << (hidden_rows % 32)).astype(np.uint32))
        payload = (1).to_bytes(4, "little") + bytes([0x04, 0x00]) + words.tobytes()
    rel = capi.DeviceRelation(ctx, n, [P.BPCHAR1, P.NUMERIC])
    cols = [(0, 1, -1, 2, 4, 0, len(flags[6])), (1, 2, -1, 1, 4, 0, -1)]
    rows, hidden = rel.load_segfile(base, 1, True, cols, [(0, payload)])
    assert (rows, hidden) == (n, len(hidden_rows))
    visible = np.ones(n, dtype=bool)
    visible[hidden_rows - 1] = False
    assert np.array_equal(rel.read_visimap(), visible)
    host = HostRelation("t", ["f", "p"], [P.BPCHAR1, P.NUMERIC], [flags[7].astype(np.uint8), price[7]], nulls=[flags[8], None],
                        visimap=np.packbits(visible, bitorder="little"))
    sc = P.SeqScan(1, [("f", P.Var(1, 1, P.BPCHAR1)), ("p", P.Var(1, 2, P.NUMERIC, 2))])
    v = _child_var(sc)
    plan = P.Agg(sc, P.AGG_HASHED, P.AGGSPLIT_SIMPLE, [1], [("f", v("f")), ("s", P.Aggref(P.AGG_SUM, v("p"))), ("n", P.Aggref(P.AGG_COUNT_STAR))],
                 num_groups=8)
    ex = capi.Executor(ctx, [rel])
    got = ex.run(plan)
    want = oracle.execute(plan, [[host]])
    assert canon(got.rows) == canon(want.rows)
    ex.close()
    # without the EOF the trailing garbage is read as a block header and refused; a missing file is named
    with pytest.raises(capi.CbgpuError):
        rel.load_segfile(base, 1, True, [(0, 1, -1, 2, 4, 0, -1)])
    with pytest.raises(capi.CbgpuError) as e:
        rel.load_segfile(base, 2, True, cols)
    assert "24576.2" in str(e.value)
    with pytest.raises(capi.CbgpuError):
        rel.load_segfile(base, 1, True, [(0, 1, -1, 2, 4, 0, len(flags[6]) + 4000)])     # EOF beyond the file
    rel.free()
