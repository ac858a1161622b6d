"""The vectors of the reference's OWN unit tests for this path (src/backend/**/test/*_test.c, cmockery), replayed against the
restatements here (not gpu).  The reference runs these against its C code; here the same inputs go through the reference's
writer / codec as compiled into oracle/_ref and come back through oracle/aocs_format.py (the restated readers that the
device kernels are tested against) or through the product's host entry points.

  access/appendonly/test/aomd_test.c:84-116                test__FormatAOSegmentFileName
  utils/misc/test/bitmap_compression_test.c:66-452         Raw / ExplicitNoCompression / ImplicitNoCompression / MultipleTypeBitmap
  utils/datumstream/test/datumstreamblock_test.c:28-165     test__DeltaCompression__Core
  access/appendonly/test/appendonly_visimap_entry_test.c:9-27   AppendOnlyVisimapEntry_GetFirstRowNum
"""
import os

import numpy as np
import pytest

from cloudberry_b200 import capi
from oracle import aocs_format as F

HERE = os.path.dirname(os.path.abspath(__file__))
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(HERE, "..", "oracle", "_ref", "libaocs_ref.so")),
                               reason="oracle/_ref not built (no /root/reference on this box)")


def test_format_ao_segment_file_name():
    base = "base/21381/123"
    # the reference's (segno, column) -> here (segno, filenum = column + 1); "no columns" (-1) is the row-oriented AO file = filenum 1
    assert capi.aocs_segfile_path(base, 0, 1) == "base/21381/123"
    assert capi.aocs_segfile_path(base, 1, 1) == "base/21381/123.1"
    assert capi.aocs_segfile_path(base, 0, 2) == "base/21381/123.128"
    assert capi.aocs_segfile_path(base, 1, 2) == "base/21381/123.129"
    assert capi.aocs_segfile_path(base, 0, 3) == "base/21381/123.256"


RAW = [0xFFFFFFFF, 0xFF00FF00, 0xFF00FF00, 0xFFFFFFFF]
IMPLICIT_NO = [0x00FFFFFF, 0xFF00FF00, 0xFFFF00FF, 0xFFFFFF00]
MULTI = [0xFFFFFFFF, 0xFF00FF00, 0xFF00FF00, 0xFFFFFFFF, 0xFFFFFFFF, 0x00000000] + [0xFFFFFFFF] * 8 + [0xFF22FF00, 0xFF11FF00]


def _offsets(words):
    return [32 * i + b for i, w in enumerate(words) for b in range(32) if (w >> b) & 1]


@needs_ref
@pytest.mark.parametrize("words,raw,size_rule", [
    (RAW, False, lambda r: 0 <= r < 16),                  # test__BitmapCompression__Raw: r < sizeof(uint32) * 4
    (RAW, True, lambda r: r == 16 + 2),                   # ..ExplicitNoCompression: r == 4 words + 2
    (IMPLICIT_NO, False, lambda r: r == 16 + 2),          # ..ImplicitNoCompression: compression does not pay -> stored raw
    (MULTI, False, lambda r: 0 <= r < 64),                # ..MultipleTypeBitmap: zero / ones / raw / repeat tokens in one stream
])
def test_bitmap_compression_vectors(words, raw, size_rule):
    payload = F.ref_visimap_entry(_offsets(words), raw=raw)
    assert size_rule(len(payload) - 4)                    # - the int32 version in front of the compressed bitmap
    got = F.visimap_entry_blocks(payload)
    assert [int(x) for x in got] == words
    # compression type bit and block count as BitmapDecompress_Init reads them (MSB first: 1 + 3 + 12 bits)
    hdr = int.from_bytes(payload[4:6], "big")
    assert hdr >> 15 == (0 if (raw or words is IMPLICIT_NO) else 1)
    assert hdr & 0xFFF == len(words)


@needs_ref
def test_bitmap_compression_no_blocks():
    """..ExplicitNoCompressionNoBlocks: an empty bitmap is the 2-byte header alone"""
    payload = F.ref_visimap_entry([], raw=True)
    assert len(payload) - 4 == 2
    assert len(F.visimap_entry_blocks(payload)) == 0


@needs_ref
def test_delta_compression_core_sequence():
    """test__DeltaCompression__Core's six int4 values: 32 stored, +1, +20, -30 as deltas, a jump beyond
    MAX_DELTA_SUPPORTED_DELTA_COMPRESSION (0x1FFFFFFF, datumstreamblock.c:2986) stored again, then a negative delta back"""
    MAXD = 0x1FFFFFFF
    vals = [32, 33, 53, 23, 23 + MAXD + 1, 63]
    raw, nblocks = F.ref_write_column("int4", vals, checksum=True, rle=2)
    assert nblocks == 1
    got, nulls = F.decode_column(raw, "int4", True)
    assert got.tolist() == vals and not nulls.any()
    # the block itself: Dense version with the delta extension, 2 physical datums (the first value and the jump), 6 delta-bitmap
    # positions of which 4 are on (the test's bitCount / bitOnCount / physical_datum_count)
    (blk, rows), = list(F.block_contents(raw, True))
    assert rows == 6
    flags = int.from_bytes(blk[2:4], "little")
    assert flags & 4
    p = 16 + (16 if flags & 2 else 0)
    d_count, d_items, d_size = [int.from_bytes(blk[p + 4 * i:p + 4 * i + 4], "little") for i in range(3)]
    assert d_count == 6 and d_items == 4
    assert int.from_bytes(blk[8:12], "little", signed=True) == 2         # physical_datum_count
    # the largest delta the format takes is still a delta; one more is not
    vals2 = [0, MAXD, 2 * MAXD + 1 - (1 << 32) if 2 * MAXD + 1 >= (1 << 31) else 2 * MAXD + 1]
    raw2, _ = F.ref_write_column("int4", vals2[:2], checksum=True, rle=2)
    got2, _ = F.decode_column(raw2, "int4", True)
    assert got2.tolist() == vals2[:2]


def test_visimap_entry_first_row_num():
    """AppendOnlyVisimapEntry_GetFirstRowNum: the entry of row r starts at r - r % 32768, beyond INT32_MAX too"""
    assert F.VISIMAP_RANGE == 32768
    for row, first in ((5, 0), (3000000000, 2999975936)):
        assert row - row % F.VISIMAP_RANGE == first
