#!/usr/bin/env python3
"""Write tests/golden/tupser_rows.npz: rows and the tuple chunk streams the REFERENCE's own SerializeTuple
(cdb/motion/tupser.c, over heap_form_minimal_tuple, access/common/heaptuple.c; compiled into oracle/_ref, driven by
oracle/ref_tupser.c) produces for them.  Runs only where /root/reference exists (`make -C oracle` first)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import aocs_format as A  # noqa: E402
from oracle import tupser as T  # noqa: E402

MODES = ["AIR", "FOB", "MAIL", "RAIL", "REG AIR", "SHIP", "TRUCK"]
WORDS = sorted(["", " ", "a", "carefully final deposits", "x" * 126, "y" * 127, "z" * 300, "quick  ", "quick"], key=lambda s: s.encode())


def cases(rng):
    n = 300
    # a Q1-like result row, a lineitem-like row with NULLs, wide rows (NULL bitmap of several bytes), strings
    yield "q1_groups", [("bpchar", 0, 1), ("bpchar", 0, 1), ("numeric", 2, 0), ("numeric", 4, 0), ("numeric", 6, 0), ("int8", 0, 0)], \
        [[chr(rng.choice([65, 78, 82])), chr(rng.choice([70, 79])), int(rng.integers(-10**15, 10**15)), int(rng.integers(0, 10**17)),
          int(rng.integers(-10**18, 10**18)), int(rng.integers(0, 10**9))] for _ in range(n)], None, 8160
    cols = [("int4", 0, 0), ("int8", 0, 0), ("date", 0, 0), ("float8", 0, 0), ("bool", 0, 0), ("numeric", 2, 0)]
    rows = [[int(rng.integers(-2**31, 2**31)), int(rng.integers(-2**62, 2**62)), int(rng.integers(-3000, 9000)), float(rng.normal(0, 1e6)),
             int(rng.integers(0, 2)), int(rng.choice([0, 1, -1, 100, 10000, 99999999, int(rng.integers(-10**12, 10**12))]))] for _ in range(n)]
    yield "mixed_nulls", cols, rows, (rng.random((n, 6)) < 0.2).astype(np.uint8).tolist(), 8160
    wide = [("int4", 0, 0), ("bool", 0, 0), ("int8", 0, 0)] * 7                      # 21 attributes: 3 bitmap bytes
    yield "wide_21_atts", wide, [[int(rng.integers(-100, 100)) if k % 3 != 1 else int(rng.integers(0, 2)) for k in range(21)] for _ in range(60)], \
        (rng.random((60, 21)) < 0.3).astype(np.uint8).tolist(), 8160
    yield "strings", [("int4", 0, 0), ("bpchar", 0, 10), ("text", 0, 0), ("int8", 0, 0)], \
        [[int(rng.integers(0, 100)), MODES[int(rng.integers(0, 7))], WORDS[int(rng.integers(0, len(WORDS)))], int(rng.integers(0, 10**12))]
         for _ in range(n)], (rng.random((n, 4)) < 0.1).astype(np.uint8).tolist(), 8160
    # tuples larger than a chunk: TC_PARTIAL_START / MID / END
    yield "chunked_small_packets", [("text", 0, 0), ("int8", 0, 0), ("text", 0, 0)], \
        [[WORDS[int(rng.integers(0, len(WORDS)))], int(rng.integers(0, 10**12)), WORDS[int(rng.integers(0, len(WORDS)))]] for _ in range(80)], None, 64
    yield "no_attributes", [], [[], [], []], None, 8160


def main():
    if A.ref_lib() is None:
        sys.exit("oracle/_ref/libaocs_ref.so missing")
    rng = np.random.default_rng(20260927)
    out = {}
    meta = []
    for name, cols, rows, nulls, max_chunk in cases(rng):
        data, nchunks = T.serialize(cols, rows, nulls, max_chunk)
        out[name + "__chunks"] = np.frombuffer(data or b"\0", dtype=np.uint8)[:len(data)]
        meta.append({"name": name, "cols": cols, "rows": rows, "nulls": nulls, "max_chunk": max_chunk, "nchunks": nchunks, "nbytes": len(data)})
        print("  %-24s %3d rows  %5d chunks  %7d bytes" % (name, len(rows), nchunks, len(data)))
    out["meta"] = np.array(json.dumps(meta))
    path = os.path.join(ROOT, "tests", "golden", "tupser_rows.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
