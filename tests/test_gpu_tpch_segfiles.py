"""From the reference's bytes to the reference's answers: the rpt_tpch tables as AOCS column files written by the REFERENCE's
block writer (tests/golden/rpt_tpch_segfiles.npz), laid out on disk under the reference's file names, read and decoded by
cb_aocs_load_segfile (CRC-32C, datum streams, numeric, character(n) dictionaries -- all on the device), then Q1 / Q3 / Q5
== the rows the reference's own regression expects (output/rpt_tpch.source, tests/golden/rpt_tpch_expected.json)."""
import os

import numpy as np
import pytest

from cloudberry_b200 import capi, tpch
from cloudberry_b200 import plan as P

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CHARS = {"c_mktsegment", "n_name", "r_name"}


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def loaded(ctx, golden, tmp_path_factory):
    d = np.load(os.path.join(HERE, "golden", "rpt_tpch_segfiles.npz"))
    root = tmp_path_factory.mktemp("base")
    host_rels, exp = golden
    rels, dicts = [], {}
    for t, (table, cols) in enumerate(tpch.SCHEMA.items()):
        base = str(root / str(16384 + t))
        segnos = sorted({int(k.split("/")[2]) for k in d.files if k.startswith(table + "/")})
        eofs = {}
        for filenum, (name, typ) in enumerate(cols, start=1):
            for segno in segnos:
                raw = bytes(d["%s/%s/%d" % (table, name, segno)])
                eofs[name, segno] = len(raw)
                with open(capi.aocs_segfile_path(base, segno, filenum), "wb") as f:
                    f.write(raw + b"\0" * 64)                      # unflushed tail of a later append: beyond the EOF

        def spec(i, name, typ, segno):
            eof = eofs[name, segno]
            if name in CHARS:
                return (i, i + 1, -1, 3, 4, 0, eof, dicts[name])
            if typ == P.NUMERIC:
                return (i, i + 1, -1, 1, 4, 0, eof)
            if typ == P.BPCHAR1:
                return (i, i + 1, -1, 2, 4, 0, eof)
            w = 8 if typ == P.INT8 else 4
            return (i, i + 1, w, 0, w, 0, eof)
        for i, (name, typ) in enumerate(cols):
            if name in CHARS:
                dicts[name] = capi.DeviceDict(ctx, max_entries=64)
                for segno in segnos:
                    capi.aocs_dict_collect_segfile(ctx, base, segno, True, spec(i, name, typ, segno))
                dicts[name].finalize()
        nrows = host_rels[t].nrows
        rel = capi.DeviceRelation(ctx, nrows, [typ for _, typ in cols])
        off = 0
        for segno in segnos:
            n, hidden = rel.load_segfile(base, segno, True, [spec(i, name, typ, segno) for i, (name, typ) in enumerate(cols)], row_offset=off)
            assert hidden == 0
            off += n
        assert off == nrows
        rels.append(rel)
    yield rels, dicts, exp, host_rels
    for r in rels:
        r.free()
    for x in dicts.values():
        x.free()


def test_columns_equal_the_csv_values(loaded):
    """every decoded column == the values parsed from the reference's csv files (dictionary columns through their texts)"""
    rels, dicts, exp, host_rels = loaded
    for rel, host, (table, cols) in zip(rels, host_rels, tpch.SCHEMA.items()):
        for i, (name, typ) in enumerate(cols):
            got, gotnull = rel.read_column(i)
            assert not gotnull.any()
            if name in CHARS:
                texts = [t.decode() for t in dicts[name].entries()]
                assert [texts[c] for c in got] == [host.dict_texts[i][c] for c in host.columns[i]], name
            else:
                assert np.array_equal(got.astype(np.int64), host.columns[i].astype(np.int64)), name


@pytest.mark.parametrize("generic", [False, True])
def test_q1_q3_q5_reference_expected(ctx, loaded, generic):
    rels, dicts, exp, host_rels = loaded
    ex = capi.Executor(ctx, rels, force_generic=generic)
    assert tpch.format_q1(ex.run(tpch.q1_plan(1)).rows) == exp["q1"]
    seg = dicts["c_mktsegment"].lookup("MACHINERY")
    assert seg >= 0
    assert tpch.format_q3(ex.run(tpch.q3_plan(seg, 1)).rows) == exp["q3"]
    reg = dicts["r_name"].lookup("AMERICA")
    names = [t.decode() for t in dicts["n_name"].entries()]
    assert tpch.format_q5(ex.run(tpch.q5_plan(reg, 1)).rows, names) == exp["q5"]
    ex.close()
