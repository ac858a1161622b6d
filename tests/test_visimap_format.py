"""pg_aovisimap entries (visimap), CPU side: the restated bitmap decoder and row lookup
(oracle/aocs_format.py) against tests/golden/aocs_visimap.npz -- entries written by the REFERENCE's Bitmap_Compress with
the set of hidden row numbers that went in -- and, where the reference library is present, against its
BitmapDecompress_Decompress."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import aocs_format as A

HERE = os.path.dirname(os.path.abspath(__file__))


def golden():
    d = np.load(os.path.join(HERE, "golden", "aocs_visimap.npz"))
    for c in d["cases"]:
        name, base, checksum = str(c).split("|")
        payload = bytes(d[name + "__payload"])
        entries = [(int(f), None if o < 0 else payload[int(o):int(o) + int(n)])
                   for f, o, n in zip(d[name + "__first"], d[name + "__off"], d[name + "__len"])]
        yield name, bytes(d[name + "__raw"]), int(checksum), entries, d[name + "__visible"]


VCASES = list(golden())


@pytest.mark.parametrize("case", VCASES, ids=[c[0] for c in VCASES])
def test_restated_lookup_matches_what_was_hidden(case):
    name, raw, checksum, entries, visible = case
    got = A.visimap_visible(raw, checksum, dict(entries))
    assert np.array_equal(got, visible)


def test_fixture_reaches_the_token_kinds():
    """zero / all-ones / raw / repeat tokens and the uncompressed type all occur in the fixture entries"""
    kinds = set()
    for name, raw, checksum, entries, visible in VCASES:
        for first, payload in entries:
            if payload is None:
                kinds.add("null")
                continue
            data = payload[4:]
            if not data[0] & 0x80:
                kinds.add("type0")
                continue
            blocks = A.visimap_entry_blocks(payload)
            if (blocks == 0).any():
                kinds.add("zero")
            if (blocks == 0xFFFFFFFF).any():
                kinds.add("ones")
            if ((blocks != 0) & (blocks != 0xFFFFFFFF)).any():
                kinds.add("raw")
            if len(blocks) > 8 and (np.diff(blocks.astype(np.int64)) == 0).sum() > 6:
                kinds.add("repeat")
    assert kinds >= {"null", "type0", "zero", "ones", "raw", "repeat"}, kinds


def test_malformed_entries():
    good = VCASES[0][3][0][1]
    with pytest.raises(ValueError):
        A.visimap_entry_blocks(b"\x02\0\0\0" + good[4:])            # version
    with pytest.raises(ValueError):
        A.visimap_entry_blocks(good[:len(good) // 2])                  # stream ends early
    with pytest.raises(ValueError):
        A.visimap_entry_blocks(b"\x01\0\0\0" + bytes([0x80, 0x01, 0x80]))     # repeat token first


@pytest.mark.skipif(A.ref_lib() is None, reason="reference library only where /root/reference exists")
def test_restated_decoder_against_reference_codec():
    L = A.ref_lib()
    rng = np.random.default_rng(1)
    for t in range(200):
        k = t % 6
        if k == 0:
            offs = []
        elif k == 1:
            offs = rng.integers(0, 32, 5)
        elif k == 2:
            offs = rng.integers(0, 32768, int(rng.integers(1, 3000)))
        elif k == 3:
            offs = np.arange(int(rng.integers(0, 20000)), int(rng.integers(20000, 32768)))
        elif k == 4:
            offs = np.concatenate([np.arange(64, 64 + int(rng.integers(1, 5)) * 32), rng.integers(0, 1000, 3)])
        else:
            offs = np.arange(0, 32768, int(rng.integers(1, 70)))
        offs = np.unique(np.asarray(offs, dtype=np.int32))
        for raw in (False, True):
            p = A.ref_visimap_entry(offs, raw=raw)
            mine = A.visimap_entry_blocks(p)
            blk = (C.c_uint32 * 1024)()
            bc = L.ref_visimap_entry_read(p, len(p), blk, 1024)
            assert bc == len(mine) and np.array_equal(np.frombuffer(blk, dtype=np.uint32)[:bc], mine)
            assert np.array_equal(np.nonzero(np.unpackbits(mine.view(np.uint8), bitorder="little"))[0], offs)
