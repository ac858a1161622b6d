"""Motion tuple serialisation through the REFERENCE's own code (oracle/_ref/libaocs_ref.so: heaptuple.c, tupser.c,
tupchunklist.c driven by oracle/ref_tupser.c).  Test infrastructure: what a Motion sender puts on the wire for a row,
what a receiver makes of a chunk stream.

Columns are (kind, dscale, n): kind in int4 / int8 / date / float8 / bool / numeric / bpchar (character(n), value = text,
blank-padded to n here as bpcharin does) / text / bytea (value = bytes: a partial aggregate's serialised state)."""
import ctypes as C

import numpy as np

from . import aocs_format as A

# kind -> (type oid, typlen, byval, align, storage)
KINDS = {"int4": (23, 4, 1, "i", "p"), "int8": (20, 8, 1, "d", "p"), "date": (1082, 4, 1, "i", "p"), "float8": (701, 8, 1, "d", "p"),
         "bool": (16, 1, 1, "c", "p"), "numeric": (1700, -1, 0, "i", "m"), "bpchar": (1042, -1, 0, "i", "x"), "text": (25, -1, 0, "i", "x"),
         "bytea": (17, -1, 0, "i", "x")}


def _lib():
    L = A.ref_lib()
    if L is None:
        return None
    L.ref_tupser_serialize.restype = C.c_int64
    L.ref_tupser_serialize.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64,
                                                                     C.POINTER(C.c_int64)]
    L.ref_tupser_deserialize.restype = C.c_int64
    L.ref_tupser_deserialize.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    return L


def _desc(cols):
    arrs = []
    for k in range(5):
        vals = [KINDS[c[0]][k] if k < 3 else ord(KINDS[c[0]][k]) for c in cols]
        arrs.append((C.c_int * max(len(cols), 1))(*vals))
    return arrs


def datum_bytes(col, v):
    """the varlena a backend would hold for the value (4-byte header)"""
    kind, dscale, n = col
    if kind == "numeric":
        return A.numeric_varlena(int(v), dscale)
    b = v if isinstance(v, bytes) else str(v).encode()
    if kind == "bpchar":
        b = b.ljust(n)
    return int((4 + len(b)) << 2).to_bytes(4, "little") + b


def serialize(cols, rows, nulls=None, max_chunk=8160):
    """rows: per row a list (ints / floats / scaled numerics / text); -> (chunk bytes, number of chunks)"""
    L = _lib()
    n = len(cols)
    vals = np.zeros((len(rows), max(n, 1)), dtype=np.int64)
    nl = np.zeros((len(rows), max(n, 1)), dtype=np.uint8)
    var = bytearray()
    for r, row in enumerate(rows):
        for a, (col, v) in enumerate(zip(cols, row)):
            if nulls is not None and nulls[r][a]:
                nl[r, a] = 1
                continue
            if KINDS[col[0]][1] == -1:
                b = datum_bytes(col, v)
                var += b"\0" * ((-len(var)) % 4)
                vals[r, a] = len(var)
                var += b
            elif col[0] == "float8":
                vals[r, a] = np.array([v], dtype=np.float64).view(np.int64)[0]
            else:
                w = KINDS[col[0]][1]
                vals[r, a] = int(v) & ((1 << (8 * w)) - 1) if w < 8 else int(v)
    vb = (C.c_ubyte * (len(var) + 8)).from_buffer_copy(bytes(var) + b"\0" * 8)
    cap = 64 + sum(len(r) for r in rows) * 16 + len(var) * 2 + len(rows) * 64 + (len(var) // max(max_chunk - 4, 1) + len(rows) + 4) * 8
    out = (C.c_ubyte * cap)()
    nch = C.c_int64()
    desc = _desc(cols)             # kept alive across the call
    k = L.ref_tupser_serialize(n, *[C.addressof(x) for x in desc], vals.ctypes.data, C.addressof(vb), nl.ctypes.data, len(rows), max_chunk,
                               C.addressof(out), cap, C.byref(nch))
    if k < 0:
        raise RuntimeError("reference SerializeTuple failed (%d): %s" % (k, L.ref_aocs_last_error().decode()))
    return bytes(out[:k]), int(nch.value)


def deserialize(cols, data, maxrows=100000):
    """chunk bytes -> (rows of raw datums: ints for by-value, bytes of the varlena (with its header) otherwise, nulls)"""
    L = _lib()
    n = len(cols)
    vals = np.zeros((maxrows, max(n, 1)), dtype=np.int64)
    nl = np.zeros((maxrows, max(n, 1)), dtype=np.uint8)
    varcap = len(data) * 2 + 64
    var = (C.c_ubyte * varcap)()
    desc = _desc(cols)
    k = L.ref_tupser_deserialize(n, *[C.addressof(x) for x in desc], data, len(data), vals.ctypes.data, nl.ctypes.data, maxrows,
                                 C.addressof(var), varcap)
    if k < 0:
        raise ValueError("reference CvtChunksToTup rejected the stream (%d): %s" % (k, L.ref_aocs_last_error().decode()))
    rows = []
    raw = bytes(var)
    for r in range(k):
        row = []
        for a, col in enumerate(cols):
            if nl[r, a]:
                row.append(None)
            elif KINDS[col[0]][1] == -1:
                p = int(vals[r, a])
                size = (raw[p] >> 1) if raw[p] & 1 else (int.from_bytes(raw[p:p + 4], "little") >> 2)
                row.append(raw[p:p + size])
            else:
                row.append(int(vals[r, a]))
        rows.append(row)
    return rows
