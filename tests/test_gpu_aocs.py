"""AOCS column files decoded ON THE DEVICE (cbgpu_aocs_decode_column) == the values the reference's own block writer
was given (tests/golden/aocs_columns.npz, see tests/test_aocs_format.py), then a query straight off decoded files."""
import numpy as np
import pytest

from cloudberry_b200 import capi
from cloudberry_b200 import plan as P
from test_aocs_format import CASES, ZCASES, ZSTDCASES

pytestmark = pytest.mark.gpu

# typname -> (relation column type, attlen, varkind, typalign)
DECODE = {"int4": (P.INT4, 4, 0, 4), "int8": (P.INT8, 8, 0, 8), "date": (P.DATE, 4, 0, 4), "float8": (P.FLOAT8, 8, 0, 8),
          "bool": (P.BOOL, 1, 0, 1), "numeric": (P.NUMERIC, -1, 1, 4), "bpchar": (P.BPCHAR1, -1, 2, 4)}


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_device_decode_matches_reference_written_values(ctx, case):
    name, typname, checksum, blocksize, dscale, nblocks, raw, values, nulls = case
    ctype, attlen, varkind, align = DECODE[typname]
    rel = capi.DeviceRelation(ctx, len(values), [ctype], dscales=[dscale])
    n = rel.load_aocs_column(0, raw, checksum, attlen, varkind, align)
    assert n == len(values)
    got, gotnull = rel.read_column(0)
    assert np.array_equal(gotnull.astype(np.uint8), nulls)
    keep = nulls == 0
    if typname == "float8":
        assert np.array_equal(got[keep].view(np.int64), values[keep].view(np.int64))
    else:
        assert np.array_equal(got[keep].astype(np.int64), values[keep])
    rel.free()


def test_refuses_what_it_does_not_decode(ctx):
    name, typname, checksum, blocksize, dscale, nblocks, raw, values, nulls = CASES[0]
    rel = capi.DeviceRelation(ctx, len(values), [P.INT4])
    bad = bytearray(raw)
    bad[3] = (bad[3] & 0x8F) | 0x20          # header kind 2 = LargeContent
    with pytest.raises(capi.CbgpuError):
        rel.load_aocs_column(0, bytes(bad), checksum, 4, 0, 4)
    with pytest.raises(capi.CbgpuError):
        rel.load_aocs_column(0, raw[:len(raw) - 24], checksum, 4, 0, 4)      # truncated file
    with pytest.raises(capi.CbgpuError):
        rel.load_aocs_column(0, raw, checksum, 8, 0, 8)                      # wrong width for the column
    rel.free()


@pytest.mark.parametrize("name", ["int4_plain", "numeric_price", "rle_float8_many_blocks_nulls", "delta_date_sorted_nulls", "int4_tiny"])
def test_checksum_failure_is_reported(ctx, name):
    """a flipped bit anywhere in a checksummed block -> CBGPU_ERR_CORRUPT from the device-side CRC-32C check
    (AppendOnlyStorageFormat_VerifyBlockChecksum / _VerifyHeaderChecksum, cdbappendonlystorageformat.c:1657-1720)"""
    case = {c[0]: c for c in CASES}[name]
    _, typname, checksum, blocksize, dscale, nblocks, raw, values, nulls = case
    assert checksum
    ctype, attlen, varkind, align = DECODE[typname]
    rel = capi.DeviceRelation(ctx, len(values), [ctype], dscales=[dscale])
    rng = np.random.default_rng(len(raw))
    spots = {16, len(raw) - 1, len(raw) // 2, 8, 12} | set(int(x) for x in rng.integers(16, len(raw), 6))
    for pos in sorted(spots):
        if pos < 8:
            continue
        bad = bytearray(raw)
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        with pytest.raises(capi.CbgpuError) as e:
            rel.load_aocs_column(0, bytes(bad), checksum, attlen, varkind, align)
        # a flip inside a header's length bits may already stop the host-side block walk (invalid file)
        assert e.value.code in (-6, -2), (pos, e.value)
    # the untouched file still loads afterwards: the error state does not stick
    assert rel.load_aocs_column(0, raw, checksum, attlen, varkind, align) == len(values)
    got, gotnull = rel.read_column(0)
    keep = nulls == 0
    if typname == "float8":
        assert np.array_equal(got[keep].view(np.int64), values[keep].view(np.int64))
    else:
        assert np.array_equal(got[keep].astype(np.int64), values[keep])
    rel.free()


@pytest.mark.parametrize("case", ZCASES + ZSTDCASES, ids=[c[0] for c in ZCASES + ZSTDCASES])
def test_device_inflates_bulk_compressed_columns(ctx, case):
    """compresstype=zlib / rle_type compresslevel 2-4 / zstd column files, written through the reference's header makers
    and a real zlib / libzstd: decompressed on the device (k_aocs_inflate / k_aocs_unzstd), then decoded like stored blocks"""
    name, typname, checksum, blocksize, dscale, nblocks, raw, values, nulls, zlevel = case
    ctype, attlen, varkind, align = DECODE[typname]
    kind = 2 if name.startswith("zstd") else 1
    rel = capi.DeviceRelation(ctx, len(values) + 5, [ctype], dscales=[dscale])
    n = rel.load_aocs_column(0, raw, checksum, attlen, varkind, align, row_offset=5, compresstype=kind)
    assert n == len(values)
    got, gotnull = rel.read_column(0, 5, 5 + n)
    assert np.array_equal(gotnull.astype(np.uint8), nulls)
    keep = nulls == 0
    if typname == "float8":
        assert np.array_equal(got[keep].view(np.int64), values[keep].view(np.int64))
    else:
        assert np.array_equal(got[keep].astype(np.int64), values[keep])
    # a column that turned out to hold no NULL keeps no null map, one with NULLs has one
    assert bool(rel.ctx.L.cbgpu_rel_has_nulls(rel.h, 0)) == bool(nulls.any())
    if "stored" not in name:
        # without the column's compresstype the compressed blocks are refused, not misread
        with pytest.raises(capi.CbgpuError) as e:
            rel.load_aocs_column(0, raw, checksum, attlen, varkind, align)
        assert e.value.code == -3
        # the wrong decompressor sees a stream that is not its own: reported, not followed
        with pytest.raises(capi.CbgpuError) as e:
            rel.load_aocs_column(0, raw, checksum, attlen, varkind, align, compresstype=3 - kind)
        assert e.value.code == -6
    rel.free()


@pytest.mark.parametrize("name", ["zlib5_numeric_price", "zlib6_float8_nulls_8k_nocrc", "zlib6_int8_bigblocks", "rle2_numeric_long_run",
                                  "rle4_float8_runs_8k_nocrc", "zstd3_numeric_price", "zstd3_int8_bigblocks"])
def test_damaged_compressed_blocks_are_reported(ctx, name):
    """flipped bits inside compressed content: with checksums the CRC catches them, without, the inflater does (bad
    Huffman code, distance before the start, wrong length, Adler-32) -- as uncompress() / gp_decompress would"""
    case = {c[0]: c for c in ZCASES + ZSTDCASES}[name]
    _, typname, checksum, blocksize, dscale, nblocks, raw, values, nulls, zlevel = case
    ctype, attlen, varkind, align = DECODE[typname]
    kind = 2 if name.startswith("zstd") else 1
    rel = capi.DeviceRelation(ctx, len(values), [ctype], dscales=[dscale])
    rng = np.random.default_rng(len(raw) + 1)
    hdr = 8 + (8 if checksum else 0) + 16
    for pos in [hdr, hdr + 1, hdr + 2, len(raw) - 9] + [int(x) for x in rng.integers(hdr + 3, len(raw) - 8, 12)]:
        bad = bytearray(raw)
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        try:
            rel.load_aocs_column(0, bytes(bad), checksum, attlen, varkind, align, compresstype=kind)
        except capi.CbgpuError as e:
            assert e.code in (-6, -2, -3), (pos, e)
            continue
        # not reported: only acceptable without checksums, when the flip hit padding or a later block's header in a way
        # that still decodes to the same values (never silently different values)
        assert not checksum, pos
        got, gotnull = rel.read_column(0)
        keep = nulls == 0
        assert np.array_equal(gotnull.astype(np.uint8), nulls) and np.array_equal(got[keep].view(np.int64), values[keep].view(np.int64)), pos
    assert rel.load_aocs_column(0, raw, checksum, attlen, varkind, align, compresstype=kind) == len(values)
    rel.free()


def test_query_over_decoded_files(ctx, oracle):
    """scan + aggregate over a relation whose columns came from reference-written column files"""
    by = {c[0]: c for c in CASES}
    price, flags = by["numeric_price"], by["bpchar1_flags"]
    n = len(price[7])
    from cloudberry_b200.relation import HostRelation
    from gpu_util import canon
    rel = capi.DeviceRelation(ctx, n, [P.BPCHAR1, P.NUMERIC])
    assert rel.load_aocs_column(0, flags[6], flags[2], -1, 2, 4) == n
    assert rel.load_aocs_column(1, price[6], price[2], -1, 1, 4) == n
    host = HostRelation("t", ["f", "p"], [P.BPCHAR1, P.NUMERIC], [flags[7].astype(np.uint8), price[7]], nulls=[flags[8], None])
    sc = P.SeqScan(1, [("f", P.Var(1, 1, P.BPCHAR1)), ("p", P.Var(1, 2, P.NUMERIC, 2))])
    from cloudberry_b200.tpch import _child_var
    v = _child_var(sc)
    plan = P.Agg(sc, P.AGG_HASHED, P.AGGSPLIT_SIMPLE, [1], [("f", v("f")), ("s", P.Aggref(P.AGG_SUM, v("p"))), ("n", P.Aggref(P.AGG_COUNT_STAR))],
                 num_groups=8)
    ex = capi.Executor(ctx, [rel])
    got = ex.run(plan)
    want = oracle.execute(plan, [[host]])
    assert canon(got.rows) == canon(want.rows) and len(want.rows) == 6     # five flags + the NULL group
    ex.close()
    rel.free()
