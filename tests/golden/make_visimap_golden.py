#!/usr/bin/env python3
"""Write tests/golden/aocs_visimap.npz: pg_aovisimap entries produced by the REFERENCE's bitmap codec
(utils/misc/bitmap_compression.c + bitstream.c compiled into oracle/_ref/libaocs_ref.so, driven by
ref_visimap_entry_write in oracle/ref_aocs.c) for column files of tests/golden/aocs_columns.npz, together with the set of
hidden row numbers that went in -- the ground truth for AppendOnlyVisimap_IsVisible, independent of any decoder here.

Runs only where /root/reference exists (`make -C oracle` first)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import aocs_format as A  # noqa: E402

RANGE = A.VISIMAP_RANGE


def main():
    if A.ref_lib() is None:
        sys.exit("oracle/_ref/libaocs_ref.so missing")
    cols = np.load(os.path.join(ROOT, "tests", "golden", "aocs_columns.npz"))
    meta = {str(c).split("|")[0]: str(c).split("|") for c in cols["cases"]}
    rng = np.random.default_rng(20260925)
    out = {}
    cases = []

    def add(name, base, firsts, hidden, null_ranges=(), drop_ranges=(), raw_ranges=()):
        """firsts: per-block firstRowNum to patch into the base file (None = as written); hidden: row numbers"""
        raw = bytearray(bytes(cols[base + "__raw"]))
        checksum = int(meta[base][2])
        blocks = A.walk_blocks_ex(bytes(raw), checksum)
        if firsts is not None:
            assert not checksum
            for b, f in zip(blocks, firsts):
                raw[b["off"] - 8:b["off"]] = int(f).to_bytes(8, "little")
            blocks = A.walk_blocks_ex(bytes(raw), checksum)
        rownums = np.concatenate([b["first"] + np.arange(b["rows"]) for b in blocks])
        hidden = np.unique(np.asarray(hidden, dtype=np.int64))
        ranges = sorted(set((rownums // RANGE * RANGE).tolist()) | set((hidden // RANGE * RANGE).tolist()))
        firsts_out, payload, offs, lens = [], b"", [], []
        for r in ranges:
            if r in drop_ranges:
                continue                            # no pg_aovisimap row for this range: all visible
            firsts_out.append(r)
            if r in null_ranges:
                offs.append(-1)
                lens.append(0)
                continue
            p = A.ref_visimap_entry(hidden[(hidden >= r) & (hidden < r + RANGE)] - r, raw=r in raw_ranges)
            offs.append(len(payload))
            lens.append(len(p))
            payload += p
        effective = np.array([h for h in hidden if (h // RANGE * RANGE) not in null_ranges and (h // RANGE * RANGE) not in drop_ranges],
                             dtype=np.int64)
        out[name + "__raw"] = np.frombuffer(bytes(raw), dtype=np.uint8)
        out[name + "__first"] = np.array(firsts_out, dtype=np.int64)
        out[name + "__off"] = np.array(offs, dtype=np.int64)
        out[name + "__len"] = np.array(lens, dtype=np.int32)
        out[name + "__payload"] = np.frombuffer(payload or b"\0", dtype=np.uint8)
        out[name + "__visible"] = ~np.isin(rownums, effective)
        cases.append("%s|%s|%d" % (name, base, checksum))
        print("  %-28s rows %6d  blocks %3d  entries %d  hidden %d" % (name, len(rownums), len(blocks), len(firsts_out), int((~out[name + "__visible"]).sum())))

    n = len(cols["rle_numeric_long_run__values"])                    # one Dense block of > 60 000 rows, row numbers 1..n
    add("long_run_scattered", "rle_numeric_long_run", None,
        np.concatenate([rng.integers(1, n + 1, 500), np.arange(10000, 13000), np.arange(40000, 40100), [1, n, 32767, 32768, 32769]]))
    add("long_run_all_hidden_range", "rle_numeric_long_run", None, np.arange(32768, 65536))
    add("plain_no_entries", "int4_plain", None, [], drop_ranges=(0,))
    add("plain_null_entry", "int4_plain", None, rng.integers(1, 20000, 50), null_ranges=(0,))
    nb = int(meta["int4_nulls_8k"][5])
    firsts = 1 + np.arange(nb) * 40000 + rng.integers(0, 5000, nb)       # gaps in the row numbers (fast sequence jumps)
    rows_per = len(cols["int4_nulls_8k__values"]) // nb + 50
    hid = np.concatenate([f + rng.integers(0, rows_per, 200) for f in firsts] + [np.arange(firsts[3], firsts[3] + 700)])
    add("gaps_many_ranges", "int4_nulls_8k", firsts, hid, null_ranges=(int(firsts[5] // RANGE * RANGE),),
        drop_ranges=(int(firsts[7] // RANGE * RANGE),), raw_ranges=(int(firsts[2] // RANGE * RANGE),))
    add("gaps_dense_patterns", "int4_nulls_8k", 1 + np.arange(nb) * 1800,
        np.concatenate([np.arange(1, 21000, 2), np.arange(5000, 5640)]))
    out["cases"] = np.array(cases)
    path = os.path.join(ROOT, "tests", "golden", "aocs_visimap.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
