"""Rows in the reference's Motion wire format (cloudberry_b200/csrc/exec/cb_tupser.c): byte-identical
to what the REFERENCE's SerializeTuple puts on the wire (tests/golden/tupser_rows.npz, written through
cdb/motion/tupser.c + access/common/heaptuple.c compiled into oracle/_ref), and readable by its CvtChunksToTup /
heap_deform_tuple where that library is present.  Host code: no GPU needed."""
import json
import os

import numpy as np
import pytest

from cloudberry_b200 import capi
from cloudberry_b200 import plan as P
from oracle import aocs_format as A
from oracle import tupser as T

HERE = os.path.dirname(os.path.abspath(__file__))
KIND = {"int4": P.INT4, "int8": P.INT8, "date": P.DATE, "float8": P.FLOAT8, "bool": P.BOOL, "numeric": P.NUMERIC}


def golden():
    d = np.load(os.path.join(HERE, "golden", "tupser_rows.npz"))
    for m in json.loads(str(d["meta"])):
        yield m["name"], [tuple(c) for c in m["cols"]], m["rows"], m["nulls"], m["max_chunk"], m["nchunks"], bytes(d[m["name"] + "__chunks"])


CASES = list(golden())


def executor_form(cols, rows, nulls):
    """(attrs, rows as the executor holds them): float8 as bits, character(1) as its byte, other strings as codes of a
    dictionary in byte order (character(n) texts without their trailing blanks, as cbgpu_dict keeps them)"""
    attrs, maps = [], []
    for a, (kind, dscale, n) in enumerate(cols):
        if kind in KIND:
            attrs.append((KIND[kind], dscale, 0))
            maps.append(None)
        elif kind == "bpchar" and n == 1:
            attrs.append((P.BPCHAR1, 0, 0))
            maps.append("c")
        else:
            vals = {(r[a].rstrip(" ") if kind == "bpchar" else r[a]).encode() for i, r in enumerate(rows) if not (nulls and nulls[i][a])}
            texts = sorted(vals)
            attrs.append((P.DICT32, 0, n if kind == "bpchar" else 0, texts))
            maps.append({t: i for i, t in enumerate(texts)})
    out = []
    for i, r in enumerate(rows):
        row = []
        for a, (kind, dscale, n) in enumerate(cols):
            if nulls and nulls[i][a]:
                row.append(0)
            elif kind == "float8":
                row.append(int(np.array([r[a]], dtype=np.float64).view(np.int64)[0]))
            elif maps[a] == "c":
                row.append(ord(r[a]))
            elif maps[a] is not None:
                row.append(maps[a][(r[a].rstrip(" ") if kind == "bpchar" else r[a]).encode()])
            else:
                row.append(int(r[a]))
        out.append(row)
    return attrs, out


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_bytes_equal_the_reference_senders(case):
    name, cols, rows, nulls, max_chunk, nchunks, want = case
    attrs, xrows = executor_form(cols, rows, nulls)
    got = capi.tupser_rows(attrs, xrows, nulls, max_chunk, end=False)
    assert got == want
    # and back: the receiving side gives the executor's values again
    back, backnull, used, ended = capi.tupser_parse(attrs, want + capi.tupser_rows(attrs, [], end=True))
    assert ended and used == len(want) + 4
    assert backnull == (nulls if nulls is not None else [[0] * len(cols)] * len(rows))
    for r, (b, x) in enumerate(zip(back, xrows)):
        for a in range(len(cols)):
            if not (nulls and nulls[r][a]):
                assert b[a] == x[a], (r, a)


@pytest.mark.skipif(A.ref_lib() is None, reason="reference library only where /root/reference exists")
def test_reference_receiver_reads_our_chunks_and_random_rows_agree():
    rng = np.random.default_rng(5)
    kinds = ["int4", "int8", "date", "float8", "bool", "numeric", "text", "bpchar"]
    pool = ["", " ", "x", "ab ", "MAIL", "REG AIR", "q" * 125, "r" * 126, "s" * 127, "t" * 128, "u" * 1000, "trailing   "]
    for t in range(80):
        natts = int(rng.integers(1, 40))
        cols = [(kinds[int(k)], int(rng.integers(0, 7)) if kinds[int(k)] == "numeric" else 0,
                 int(rng.choice([1, 10, 25])) if kinds[int(k)] == "bpchar" else 0) for k in rng.integers(0, 8, natts)]
        nrows = int(rng.integers(1, 30))
        rows = []
        for _ in range(nrows):
            row = []
            for kind, ds, _n in cols:
                if kind == "float8":
                    row.append(float(rng.normal(0, 1e9)))
                elif kind == "bool":
                    row.append(int(rng.integers(0, 2)))
                elif kind == "int4" or kind == "date":
                    row.append(int(rng.integers(-2**31, 2**31)))
                elif kind == "numeric":
                    row.append(int(rng.choice([0, 1, -1, 10**ds, -10**(ds + 3), int(rng.integers(-2**62, 2**62)), int(rng.integers(-10**9, 10**9))])))
                elif kind == "text":
                    row.append(pool[int(rng.integers(0, len(pool)))])
                elif kind == "bpchar":
                    row.append(pool[int(rng.integers(0, 6))][:_n].rstrip(" ") if _n > 1 else chr(int(rng.integers(65, 91))))
                else:
                    row.append(int(rng.integers(-2**63, 2**63 - 1)))
            rows.append(row)
        nulls = (rng.random((nrows, natts)) < (0.0 if t % 3 == 0 else 0.25)).astype(np.uint8).tolist()
        max_chunk = int(rng.choice([8160, 40, 100, 1000]))
        want, nch = T.serialize(cols, rows, nulls, max_chunk)
        attrs, xrows = executor_form(cols, rows, nulls)
        got = capi.tupser_rows(attrs, xrows, nulls, max_chunk, end=False)
        assert got == want, (t, cols)
        ref_rows = T.deserialize(cols, got)
        for r in range(nrows):
            for a, (kind, ds, _n) in enumerate(cols):
                if nulls[r][a]:
                    assert ref_rows[r][a] is None
                elif kind == "numeric":
                    assert A.numeric_from_bytes(ref_rows[r][a][1:] if ref_rows[r][a][0] & 1 else ref_rows[r][a][4:], ds) == rows[r][a]
                elif kind == "float8":
                    assert ref_rows[r][a] == xrows[r][a]
                elif kind in ("text", "bpchar"):
                    body = ref_rows[r][a][1:] if ref_rows[r][a][0] & 1 else ref_rows[r][a][4:]
                    assert body == (rows[r][a].ljust(_n) if kind == "bpchar" else rows[r][a]).encode()
                else:
                    assert ref_rows[r][a] == rows[r][a]


def test_partial_buffers_and_malformed_streams():
    name, cols, rows, nulls, max_chunk, nchunks, data = [c for c in CASES if c[0] == "chunked_small_packets"][0]
    attrs, xrows = executor_form(cols, rows, nulls)
    # a buffer that ends inside a tuple: the rows before it come out, the rest waits for more bytes
    for cut in (0, 3, 10, 70, len(data) // 2, len(data) - 1):
        back, _, used, ended = capi.tupser_parse(attrs, data[:cut])
        assert not ended and used <= cut and back == xrows[:len(back)]
        more, _, used2, _ = capi.tupser_parse(attrs, data[used:])
        assert back + more == xrows
    bad = bytearray(data)
    bad[2] = 2                                  # first chunk claims to be TC_PARTIAL_MID
    with pytest.raises(capi.CbgpuError):
        capi.tupser_parse(attrs, bytes(bad))
    bad = bytearray(data)
    bad[4] ^= 0x10                              # the length word in front of the tuple body
    with pytest.raises(capi.CbgpuError):
        capi.tupser_parse(attrs, bytes(bad))
    # a string the receiving dictionary does not hold
    name, cols, rows, nulls, max_chunk, nchunks, data = [c for c in CASES if c[0] == "strings"][0]
    attrs, xrows = executor_form(cols, rows, nulls)
    poorer = [attrs[0], (attrs[1][0], 0, attrs[1][2], attrs[1][3][:2]), attrs[2], attrs[3]]
    with pytest.raises(capi.CbgpuError):
        capi.tupser_parse(poorer, data)
    # the direct-buffer form of a row without attributes (TC_EMPTY) is read as a row too
    assert capi.tupser_parse([], bytes([0, 0, 5, 0, 0, 0, 4, 0]))[0] == [[]]
