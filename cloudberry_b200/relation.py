"""Host-side description of a decoded AOCS relation: fixed-width column arrays.

This is the form the loader hands to the device (`cbgpu_rel_load_column`): what
`aocs_getnext` (access/aocs/aocsam.c:1418) would produce per projected column, but a whole column at
a time: numeric(15,2) as int64 scaled by 10^dscale, date as int32 days since 2000-01-01, char(1) as
the byte, bpchar/varchar as dictionary codes ("data layout in HBM").
"""
import numpy as np

from . import plan as P

NP_DTYPE = {P.INT4: np.int32, P.DATE: np.int32, P.DICT32: np.int32, P.INT8: np.int64, P.NUMERIC: np.int64,
            P.FLOAT8: np.float64, P.BPCHAR1: np.uint8, P.DICT8: np.uint8, P.BOOL: np.uint8}


class HostRelation:
    def __init__(self, name, names, types, columns, dscales=None, nulls=None, visimap=None, dict_texts=None,
                 dict_hashes=None):
        assert len(names) == len(types) == len(columns)
        self.name = name
        self.names = list(names)
        self.types = list(types)
        self.dscales = list(dscales) if dscales is not None else [2 if t == P.NUMERIC else 0 for t in types]
        self.columns = [np.ascontiguousarray(c, dtype=NP_DTYPE[t]) for c, t in zip(columns, types)]
        self.nrows = int(self.columns[0].shape[0]) if self.columns else 0
        for c in self.columns:
            assert c.shape[0] == self.nrows
        self.nulls = list(nulls) if nulls is not None else [None] * len(names)
        self.visimap = visimap            # packed bits, 1 = visible (appendonly_visimap.c:198)
        self.dict_texts = list(dict_texts) if dict_texts is not None else [None] * len(names)
        self.dict_hashes = list(dict_hashes) if dict_hashes is not None else [None] * len(names)

    def attno(self, colname):
        return self.names.index(colname) + 1

    def var(self, colname):
        """scan-level Var for a column (varno filled by the SeqScan's scanrelid at plan build)."""
        i = self.names.index(colname)
        return i + 1, self.types[i], self.dscales[i]

    def set_dict_hashes(self, hashfn):
        """Per-code hashbpchar() values (utils/adt/varchar.c:981) so dictionary columns can be
        hash keys with the reference's hash values."""
        for i, texts in enumerate(self.dict_texts):
            if texts is not None:
                self.dict_hashes[i] = np.array([hashfn(t) for t in texts], dtype=np.uint32)
        return self

    def take(self, idx):
        """Row subset (used to form per-segment shards)."""
        return HostRelation(self.name, self.names, self.types, [c[idx] for c in self.columns], self.dscales,
                            [None if n is None else n[idx] for n in self.nulls], None, self.dict_texts,
                            self.dict_hashes)

    def nbytes(self):
        return sum(c.nbytes for c in self.columns)
