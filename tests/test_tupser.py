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


@pytest.mark.skipif(A.ref_lib() is None, reason="reference library only where /root/reference exists")
def test_partial_aggregate_states_travel_as_the_reference_serialises_them():
    """a Partial Aggregate's output row (group key, sum(numeric) state, avg(int8) state, count) between a device stage and a CPU
    stage: the state columns are bytea on the wire (AGGSPLIT_INITIAL_SERIAL), (N, sum) in the executor.  Our chunks equal the
    reference's SerializeTuple over the same bytea values, the reference's receiver reads our chunks, and our receiver turns the
    reference's chunks back into states - also when the sender's sum has a smaller display scale than the receiver expects."""
    import ctypes as C
    E = capi.ex()
    E.cb_numeric_avg_serialize.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_int32]
    E.cb_int8_avg_serialize.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int32]

    def split(v):
        u = v & ((1 << 128) - 1)
        lo, hi = u & ((1 << 64) - 1), u >> 64
        return (lo - (1 << 64) if lo >= 1 << 63 else lo), (hi - (1 << 64) if hi >= 1 << 63 else hi)

    def ser(kind, n, v, ds=0):
        buf = C.create_string_buffer(256)
        lo, hi = split(v)
        k = E.cb_numeric_avg_serialize(n, lo, hi, ds, buf, 256) if kind == "numeric" else E.cb_int8_avg_serialize(n, lo, hi, buf, 256)
        assert k > 0
        return buf.raw[:k]
    rng = np.random.default_rng(9)
    cols = [("int4", 0, 0), ("bytea", 0, 0), ("bytea", 0, 0), ("int8", 0, 0)]
    attrs = [(P.INT4, 0, 0), (capi.TUPSER_STATE_NUMERIC, 4, 0), (capi.TUPSER_STATE_INT8, 0, 0), (P.INT8, 0, 0)]
    states, ref_rows, our_rows = [], [], []
    for r in range(200):
        n = int(rng.integers(1, 10**9))
        vnum = int(rng.integers(-2**62, 2**62)) * int(rng.choice([1, 1, 10**6, 2**30]))       # some beyond 64 bits
        vint = int(rng.integers(-2**62, 2**62)) * int(rng.choice([1, 3, 2**20]))
        states.append(((n, *split(vnum)), (n, *split(vint))))
        ref_rows.append([r, ser("numeric", n, vnum, 4), ser("int8", n, vint), n])
        our_rows.append([r, (n, *split(vnum)), (n, *split(vint)), n])
    for max_chunk in (8160, 64):
        want, _ = T.serialize(cols, ref_rows, None, max_chunk)
        got = capi.tupser_rows(attrs, our_rows, None, max_chunk, end=False)
        assert got == want
        back, backnull, used, ended = capi.tupser_parse(attrs, want + capi.tupser_rows(attrs, [], end=True))
        assert ended and [tuple(b[1:3]) for b in back] == states and [b[0] for b in back] == list(range(200))
        refback = T.deserialize(cols, got)
        for r in range(200):
            for a in (1, 2):
                body = refback[r][a][1:] if refback[r][a][0] & 1 else refback[r][a][4:]
                assert body == ref_rows[r][a]
    # a CPU partial stage whose inputs all had display scale 2 feeding a receiver that keeps scale 4: the sum is rescaled exactly
    low = [[0, ser("numeric", 5, 12345, 2), ser("int8", 5, 7), 5]]
    data, _ = T.serialize(cols, low, None, 8160)
    back, _, _, _ = capi.tupser_parse(attrs, data + capi.tupser_rows(attrs, [], end=True))
    assert back[0][1] == (5, 1234500, 0)
    # damaged state bytes are a malformed stream, not a crash or a wrong state
    for damage in ("nan", "short"):
        st = bytearray(ser("numeric", 5, 12345, 2))
        if damage == "nan":
            st[12:14] = b"\xc0\x00"             # NUMERIC_NAN in the sum's sign field: a NaN went into the CPU stage
        else:
            st = st[:-3]
        data, _ = T.serialize(cols, [[0, bytes(st), ser("int8", 5, 7), 5]], None, 8160)
        with pytest.raises(capi.CbgpuError):
            capi.tupser_parse(attrs, data + capi.tupser_rows(attrs, [], end=True))
