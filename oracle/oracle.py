"""ctypes loader for the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Never imported by the product package cloudberry_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from cloudberry_b200 import plan as P

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, f) for f in ("oracle.c", "oracle.h", "pg_hash.h")] + \
           [os.path.join(HERE, "..", "include", "cb_plan.h")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale or (os.path.isdir("/root/reference") and
                          not os.path.exists(os.path.join(HERE, "_ref", "libpg_hashfn.so"))):
        subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    return so


class OraRel(C.Structure):
    _fields_ = [("nrows", C.c_int64), ("ncols", C.c_int32), ("types", C.POINTER(C.c_int32)),
                ("dscales", C.POINTER(C.c_int32)), ("data", C.POINTER(C.c_void_p)),
                ("nulls", C.POINTER(C.c_void_p)), ("visimap", C.c_void_p),
                ("dict_hash", C.POINTER(C.c_void_p))]


class OraSegment(C.Structure):
    _fields_ = [("nrels", C.c_int32), ("rels", C.POINTER(C.POINTER(OraRel)))]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.ora_execute.restype = C.c_void_p
        L.ora_execute.argtypes = [C.POINTER(P.CbPlan), C.POINTER(OraSegment), C.c_int32, C.c_int32]
        L.ora_last_error.restype = C.c_char_p
        L.ora_result_nrows.restype = C.c_int64
        L.ora_result_nrows.argtypes = [C.c_void_p]
        L.ora_result_ncols.restype = C.c_int32
        L.ora_result_ncols.argtypes = [C.c_void_p]
        L.ora_result_type.restype = C.c_int32
        L.ora_result_type.argtypes = [C.c_void_p, C.c_int32]
        L.ora_result_segment.restype = C.c_int32
        L.ora_result_segment.argtypes = [C.c_void_p, C.c_int64]
        L.ora_result_isnull.restype = C.c_int
        L.ora_result_isnull.argtypes = [C.c_void_p, C.c_int64, C.c_int32]
        L.ora_result_int64.restype = C.c_int64
        L.ora_result_int64.argtypes = [C.c_void_p, C.c_int64, C.c_int32]
        L.ora_result_float8.restype = C.c_double
        L.ora_result_float8.argtypes = [C.c_void_p, C.c_int64, C.c_int32]
        L.ora_result_text.restype = C.c_char_p
        L.ora_result_text.argtypes = [C.c_void_p, C.c_int64, C.c_int32]
        L.ora_result_state_n.restype = C.c_int64
        L.ora_result_state_n.argtypes = [C.c_void_p, C.c_int64, C.c_int32]
        L.ora_result_state_sum.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.ora_result_free.argtypes = [C.c_void_p]
        L.ora_hash_datum.restype = C.c_uint32
        L.ora_hash_datum.argtypes = [C.c_int32, C.c_int64]
        L.ora_hashbpchar_text.restype = C.c_uint32
        L.ora_hashbpchar_text.argtypes = [C.c_char_p, C.c_int32]
        L.ora_cdbhash_segment.restype = C.c_int32
        L.ora_cdbhash_segment.argtypes = [C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_uint8), C.c_int32, C.c_int32]
        L.ora_numeric_sum_text.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_char_p, C.c_int32]
        L.ora_numeric_avg_text.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_int64, C.c_char_p, C.c_int32]
        _LIB = L
    return _LIB


def ref_hash_lib():
    """The reference's own src/common/hashfn.c, compiled (oracle/_ref); None when not built."""
    so = os.path.join(HERE, "_ref", "libpg_hashfn.so")
    if not os.path.exists(so):
        return None
    L = C.CDLL(so)
    L.hash_bytes.restype = C.c_uint32
    L.hash_bytes.argtypes = [C.c_char_p, C.c_int]
    L.hash_bytes_uint32.restype = C.c_uint32
    L.hash_bytes_uint32.argtypes = [C.c_uint32]
    L.ref_murmurhash32.restype = C.c_uint32
    L.ref_murmurhash32.argtypes = [C.c_uint32]
    return L


def hashbpchar(text):
    b = text.encode() if isinstance(text, str) else text
    return int(lib().ora_hashbpchar_text(b, len(b)))


def make_rel(rel):
    """rel: cloudberry_b200.relation.HostRelation-like object with .columns (list of numpy arrays),
    .types, .dscales, optional .nulls (list of uint8 arrays or None), .visimap (packed bits) and
    .dict_hashes (list of uint32 arrays or None).  Returns (OraRel, keepalive)."""
    n = len(rel.columns)
    keep = []
    types = (C.c_int32 * n)(*rel.types)
    dsc = (C.c_int32 * n)(*rel.dscales)
    data = (C.c_void_p * n)()
    nulls = (C.c_void_p * n)()
    dh = (C.c_void_p * n)()
    for i, col in enumerate(rel.columns):
        a = np.ascontiguousarray(col)
        keep.append(a)
        data[i] = a.ctypes.data
        nl = rel.nulls[i] if getattr(rel, "nulls", None) else None
        if nl is not None:
            nl = np.ascontiguousarray(nl, dtype=np.uint8)
            keep.append(nl)
            nulls[i] = nl.ctypes.data
        d = rel.dict_hashes[i] if getattr(rel, "dict_hashes", None) else None
        if d is not None:
            d = np.ascontiguousarray(d, dtype=np.uint32)
            keep.append(d)
            dh[i] = d.ctypes.data
    r = OraRel(nrows=rel.nrows, ncols=n, types=types, dscales=dsc, data=data, nulls=nulls, dict_hash=dh)
    vm = getattr(rel, "visimap", None)
    if vm is not None:
        vm = np.ascontiguousarray(vm, dtype=np.uint8)
        keep.append(vm)
        r.visimap = vm.ctypes.data
    keep += [types, dsc, data, nulls, dh]
    return r, keep


class Result:
    """Rows of an oracle run; numeric values as the exact text numeric_out() prints."""

    def __init__(self, handle):
        L = lib()
        self.nrows = L.ora_result_nrows(handle)
        self.ncols = L.ora_result_ncols(handle)
        self.types = [L.ora_result_type(handle, c) for c in range(self.ncols)]
        self.rows = []
        self.segments = []
        self.states = []
        for r in range(self.nrows):
            row = []
            st = []
            for c in range(self.ncols):
                if L.ora_result_isnull(handle, r, c):
                    row.append(None)
                    st.append(None)
                    continue
                t = self.types[c]
                lo, hi = C.c_int64(), C.c_int64()
                L.ora_result_state_sum(handle, r, c, C.byref(lo), C.byref(hi))
                st.append((L.ora_result_state_n(handle, r, c), (hi.value << 64) | (lo.value & (2 ** 64 - 1))))
                if t == P.FLOAT8:
                    row.append(L.ora_result_float8(handle, r, c))
                elif t in (P.NUMERIC, P.NUMERIC128):
                    row.append(L.ora_result_text(handle, r, c).decode())
                else:
                    row.append(L.ora_result_int64(handle, r, c))
            self.rows.append(row)
            self.states.append(st)
            self.segments.append(L.ora_result_segment(handle, r))
        L.ora_result_free(handle)


def execute(plan_node, segments, nthreads=1):
    """segments: list (one per segment) of lists of host relations (range table order)."""
    L = lib()
    keep = []
    segs = (OraSegment * len(segments))()
    for s, rels in enumerate(segments):
        arr = (C.POINTER(OraRel) * max(len(rels), 1))()
        for i, rel in enumerate(rels):
            r, k = make_rel(rel)
            keep += [r, k]
            arr[i] = C.pointer(r)
        segs[s].nrels = len(rels)
        segs[s].rels = arr
        keep.append(arr)
    h = L.ora_execute(P.plan_ptr(plan_node), segs, len(segments), nthreads)
    if not h:
        raise RuntimeError("oracle: " + L.ora_last_error().decode())
    return Result(h)


# ------------------------------------------------------------------------------------------------
# Q1 through the reference's own per-row functions (oracle/ref_q1.c in oracle/_ref/libexec_ref.so)
# ------------------------------------------------------------------------------------------------
_REFEXEC = None


def ref_exec_lib():
    """oracle/_ref/libexec_ref.so (the reference's hashfunc.c / varchar.c / cdbhash.c / numeric.c / datumstreamblock.c compiled
    where they lie + oracle/ref_exec.c, ref_q1.c); None when not built."""
    global _REFEXEC
    if _REFEXEC is None:
        so = os.path.join(HERE, "_ref", "libexec_ref.so")
        if not os.path.exists(so):
            return None
        R = C.CDLL(so)
        R.ref_exec_last_error.restype = C.c_char_p
        R.ref_q1_load.restype = C.c_void_p
        R.ref_q1_load.argtypes = [C.c_int64] + [C.c_void_p] * 7 + [C.c_int, C.c_int]
        R.ref_q1_run.argtypes = [C.c_void_p, C.c_int32, C.c_char_p, C.c_int, C.POINTER(C.c_int64)]
        R.ref_q1_file_bytes.restype = C.c_int64
        R.ref_q1_file_bytes.argtypes = [C.c_void_p]
        R.ref_q1_free.argtypes = [C.c_void_p]
        _REFEXEC = R
    return _REFEXEC


class RefQ1:
    """lineitem's Q1 columns as reference-written AOCS column files in memory; run() = the reference's scan cursor, numeric
    expressions, grouping hash and sum / avg transition + final functions over them, row at a time (see ref_q1.c's header for
    what is reference code and what is glue)."""

    def __init__(self, lineitem, checksum=True, blocksize=32768):
        import numpy as np
        R = ref_exec_lib()
        if R is None:
            raise RuntimeError("oracle/_ref/libexec_ref.so is not built")
        col = lambda n: np.ascontiguousarray(lineitem.columns[lineitem.attno(n) - 1])
        dec = [col(n).astype(np.int64) for n in ("l_quantity", "l_extendedprice", "l_discount", "l_tax")]
        sd = col("l_shipdate").astype(np.int32)
        rf = col("l_returnflag").astype(np.uint8)
        ls = col("l_linestatus").astype(np.uint8)
        self.nrows = int(len(sd))
        self.h = R.ref_q1_load(self.nrows, *[c.ctypes.data for c in dec], sd.ctypes.data, rf.ctypes.data, ls.ctypes.data,
                               1 if checksum else 0, blocksize)
        if not self.h:
            raise RuntimeError("ref_q1_load: " + R.ref_exec_last_error().decode())
        self.file_bytes = int(R.ref_q1_file_bytes(self.h))

    def run(self, cutoff):
        """-> (rows as lists of text in the regression output's column order, rows that passed the qual)"""
        R = ref_exec_lib()
        out = C.create_string_buffer(1 << 16)
        passed = C.c_int64()
        ng = R.ref_q1_run(self.h, cutoff, out, 1 << 16, C.byref(passed))
        if ng < 0:
            raise RuntimeError("ref_q1_run: " + R.ref_exec_last_error().decode())
        return [ln.split("|") for ln in out.value.decode().splitlines()], int(passed.value)

    def free(self):
        if self.h:
            ref_exec_lib().ref_q1_free(self.h)
            self.h = None
