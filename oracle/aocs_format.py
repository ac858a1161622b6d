"""AOCS column-file format: CPU restatement of the READER (test infrastructure) + access to the reference's WRITER.

TEST INFRASTRUCTURE ONLY (tests/, tests/golden/make_aocs_golden.py).

Reader restated here, in numpy / plain Python, from:
  storage block header    include/cdb/cdbappendonlystorage_int.h:64-147 (AOSmallContentHeader bit fields),
                          cdb/cdbappendonlystorageformat.c:81-113 (header length: 8 + 2 x CRC + firstRowNum),
                          include/cdb/cdbappendonlystorage.h:37-42 (content rounded up to 8 bytes)
  datum stream block      include/utils/datumstreamblock.h:74-83 (DatumStreamBlock_Orig), flags :202-208,
                          utils/datumstream/datumstreamblock.c:153-354 (GetReadyOrig: header, MAXALIGNed NULL bitmap,
                          MAXALIGNed datum area), datumstreamblock.h:1442-1540 (AdvanceOrig: NULLs take no datum space;
                          fixed width: += datumlen; varlena: += VARSIZE_ANY, then skip zero pad bytes to typalign)
  dense blocks (rle_type) include/utils/datumstreamblock.h:86-170 (Dense header, RLE extension), :617-730 (repeat count
                          codec), :1725-1960 (AdvanceDense: a set compress bit = the datum repeats `count` more times; the
                          NULL bitmap has one bit per non-repeated position), datumstreamblock.c:627-1010 (GetReadyDense);
                          NonBulkDenseContent storage header include/cdb/cdbappendonlystorage_int.h:259-325
  varlena headers         include/postgres.h VARATT_IS_1B / VARSIZE_1B / VARSIZE_4B (little endian)
  numeric                 include/utils/numeric.h:103-189 (short / long headers, base-10000 digits, weight, dscale)
Pinned by tests/test_aocs_format.py against column files written by the reference's own code (oracle/ref_aocs.c).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# (typid, attlen, byval, align, storage) as pg_type has them
TYPEINFO = {
    "int4": (23, 4, 1, "i", "p"), "int8": (20, 8, 1, "d", "p"), "date": (1082, 4, 1, "i", "p"),
    "float8": (701, 8, 1, "d", "p"), "bool": (16, 1, 1, "c", "p"), "int2": (21, 2, 1, "s", "p"),
    "numeric": (1700, -1, 0, "i", "m"), "bpchar": (1042, -1, 0, "i", "x"),
    # strings kept whole (dictionary columns): character(n) values arrive blank-padded to n, as bpcharin leaves them
    "bpchars": (1042, -1, 0, "i", "x"), "varchar": (1043, -1, 0, "i", "x"), "text": (25, -1, 0, "i", "x"),
}
STRING_TYPES = ("bpchars", "varchar", "text")
NBASE = 10000


# ------------------------------------------------------------------------------------------------
# numeric <-> scaled integer
# ------------------------------------------------------------------------------------------------
def numeric_digits(scaled, dscale):
    """(sign, weight, digits) of scaled * 10^-dscale in base 10000 with leading / trailing zero digits stripped
    (make_result / strip_var, utils/adt/numeric.c)."""
    sign = scaled < 0
    v = -scaled if sign else scaled
    groups_after = (dscale + 3) // 4
    v *= 10 ** (4 * groups_after - dscale)
    digits = []
    while v:
        digits.append(v % NBASE)
        v //= NBASE
    digits.reverse()
    weight = len(digits) - groups_after - 1
    while digits and digits[0] == 0:
        digits.pop(0)
        weight -= 1
    while digits and digits[-1] == 0:
        digits.pop()
    if not digits:
        weight, sign = 0, False
    return sign, weight, digits


def numeric_varlena(scaled, dscale):
    """4-byte-header varlena of a numeric datum as numeric_in would hand it to the insert path: short numeric
    header when dscale <= 63 and -64 <= weight <= 63 (NUMERIC_CAN_BE_SHORT), else the long header."""
    sign, weight, digits = numeric_digits(int(scaled), dscale)
    if dscale <= 0x3F and -64 <= weight <= 63:
        hdr = 0x8000 | (0x2000 if sign else 0) | (dscale << 7) | (0x0040 if weight < 0 else 0) | (weight & 0x003F)
        body = int(hdr).to_bytes(2, "little")
    else:
        body = int((0x4000 if sign else 0) | (dscale & 0x3FFF)).to_bytes(2, "little") + int(weight & 0xFFFF).to_bytes(2, "little")
    body += b"".join(int(d).to_bytes(2, "little") for d in digits)
    total = 4 + len(body)
    return int(total << 2).to_bytes(4, "little") + body


def numeric_from_bytes(data, dscale_out):
    """numeric datum body (after the varlena header) -> integer scaled by 10^dscale_out (exact, else ValueError)"""
    h = int.from_bytes(data[0:2], "little")
    if (h & 0xC000) == 0x8000:
        sign = bool(h & 0x2000)
        weight = h & 0x003F
        if h & 0x0040:
            weight -= 64
        off = 2
    elif (h & 0xC000) == 0xC000:
        raise ValueError("NaN / infinity")
    else:
        sign = (h & 0xC000) == 0x4000
        weight = int.from_bytes(data[2:4], "little", signed=True)
        off = 4
    nd = (len(data) - off) // 2
    acc = 0
    for i in range(nd):
        acc = acc * NBASE + int.from_bytes(data[off + 2 * i:off + 2 * i + 2], "little")
    e10 = 4 * (weight - nd + 1) + dscale_out
    if nd == 0:
        return 0
    if e10 >= 0:
        acc *= 10 ** e10
    else:
        q, r = divmod(acc, 10 ** (-e10))
        if r:
            raise ValueError("numeric has more fractional digits than the column's scale")
        acc = q
    return -acc if sign else acc


def bpchar_varlena(text):
    b = text if isinstance(text, bytes) else text.encode()
    return int((4 + len(b)) << 2).to_bytes(4, "little") + b


# ------------------------------------------------------------------------------------------------
# the reference's writer (oracle/_ref/libaocs_ref.so, built from the reference's sources where present)
# ------------------------------------------------------------------------------------------------
_REF = None


def ref_lib():
    global _REF
    if _REF is None:
        so = os.path.join(HERE, "_ref", "libaocs_ref.so")
        if not os.path.exists(so):
            return None
        L = C.CDLL(so)
        L.ref_aocs_write_column.restype = C.c_int64
        L.ref_aocs_write_column.argtypes = [C.c_int] * 7 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                                            C.POINTER(C.c_int64)]
        L.ref_aocs_write_column_ex.restype = C.c_int64
        L.ref_aocs_write_column_ex.argtypes = [C.c_int] * 8 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                                               C.POINTER(C.c_int64)]
        L.ref_aocs_write_column_z.restype = C.c_int64
        L.ref_aocs_write_column_z.argtypes = [C.c_int] * 9 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                                              C.POINTER(C.c_int64)]
        L.ref_aocs_write_column_cb.restype = C.c_int64
        L.ref_aocs_write_column_cb.argtypes = [C.c_int] * 8 + [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                                               C.POINTER(C.c_int64)]
        L.ref_aocs_last_error.restype = C.c_char_p
        L.ref_numeric_inspect.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                          C.c_void_p, C.c_int]
        L.ref_aocs_block_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                          C.POINTER(C.c_int64), C.POINTER(C.c_int)]
        L.ref_aocs_verify_block.argtypes = [C.c_void_p, C.c_int]
        L.ref_visimap_entry_write.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.ref_visimap_entry_read.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_int]
        _REF = L
    return _REF


COMPRESS_CB = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_ubyte), C.c_int, C.POINTER(C.c_ubyte), C.c_int)


def zstd_compressor(level):
    """what the reference's zstd_compress does (gpcontrib/zstd/zstd_compression.c:104-140: ZSTD_compressCCtx at the
    column's compresslevel; 'destination too small' reported as src_sz), with the libzstd bundled in pyarrow"""
    import pyarrow as pa
    codec = pa.Codec("zstd", compression_level=level)

    def cb(src, srclen, dst, dstcap):
        z = codec.compress(C.string_at(src, srclen), asbytes=True)
        if len(z) > dstcap:
            return srclen
        C.memmove(dst, z, len(z))
        return len(z)
    return COMPRESS_CB(cb)


def ref_write_column(typname, values, nulls=None, checksum=True, blocksize=32768, dscale=0, rle=False, zlevel=0, compressor=None):
    # zlevel: zlib level of the storage layer's bulk compression (compresstype=zlib compresslevel=zlevel; with rle:
    # rle_type compresslevel 2 / 3 / 4 = zlevel 1 / 5 / 9), 0 = none
    # rle: False / 0 plain, True / 1 rle_type, 2 rle_type with delta range encoding (what the reference picks for
    # int4 / int8 / date columns under rle_type)
    """Column file bytes as the reference's insert path writes them.  values: ints / floats (by-value types), scaled
    ints (numeric) or str (bpchar)."""
    L = ref_lib()
    typid, attlen, byval, align, storage = TYPEINFO[typname]
    n = len(values)
    nl = None if nulls is None else np.ascontiguousarray(nulls, dtype=np.uint8)
    varbuf = b""
    if attlen == -1:
        offs = np.zeros(n, dtype=np.int64)
        parts = []
        pos = 0
        for i, v in enumerate(values):
            if nl is not None and nl[i]:
                continue
            b = numeric_varlena(v, dscale) if typname == "numeric" else bpchar_varlena(v)
            b += b"\0" * ((-len(b)) % 4)            # datums handed in are int-aligned palloc chunks
            offs[i] = pos
            parts.append(b)
            pos += len(b)
        varbuf = b"".join(parts) + b"\0" * 8
        vals = offs
    elif typname == "float8":
        vals = np.ascontiguousarray(values, dtype=np.float64).view(np.int64)
    else:
        vals = np.ascontiguousarray(values, dtype=np.int64)
        if attlen < 8:
            vals = vals & ((1 << (8 * attlen)) - 1)  # a by-value Datum holds the zero-extended low bytes
    cap = n * 24 + (n // 100 + 4) * 64 + 4096 + 2 * len(varbuf)
    out = (C.c_ubyte * cap)()
    nb = C.c_int64()
    vb = (C.c_ubyte * max(len(varbuf), 1)).from_buffer_copy(varbuf or b"\0")
    if compressor is not None:
        r = L.ref_aocs_write_column_cb(typid, attlen, byval, ord(align), ord(storage), 1 if checksum else 0, blocksize, int(rle),
                                       C.cast(compressor, C.c_void_p), vals.ctypes.data, C.addressof(vb),
                                       nl.ctypes.data if nl is not None else None, n, C.addressof(out), cap, C.byref(nb))
    else:
        r = L.ref_aocs_write_column_z(typid, attlen, byval, ord(align), ord(storage), 1 if checksum else 0, blocksize, int(rle), int(zlevel),
                                      vals.ctypes.data, C.addressof(vb), nl.ctypes.data if nl is not None else None, n,
                                      C.addressof(out), cap, C.byref(nb))
    if r < 0:
        raise RuntimeError("reference writer: " + L.ref_aocs_last_error().decode())
    return bytes(out[:r]), int(nb.value)


# ------------------------------------------------------------------------------------------------
# the restated reader
# ------------------------------------------------------------------------------------------------
_CRC32C_TABLE = []


def crc32c_raw(data, crc=0xFFFFFFFF):
    """CRC-32C (reflected 0x82F63B78) WITHOUT the final inversion: INIT_CRC32C + COMP_CRC32C only, which is what
    append-only block headers store (cdbappendonlystorageformat.c:41-46, 71-76; src/port/pg_crc32c_sb8.c)."""
    if not _CRC32C_TABLE:
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ (0x82F63B78 if c & 1 else 0)
            _CRC32C_TABLE.append(c)
    t = _CRC32C_TABLE
    for b in bytes(data):
        crc = t[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc


def verify_block_checksums(raw, pos, overall):
    """AppendOnlyStorageFormat_VerifyHeaderChecksum / _VerifyBlockChecksum (cdbappendonlystorageformat.c:1657-1720):
    header checksum (bytes 12..15) covers bytes [0,12); block checksum (bytes 8..11) covers [16, overall)."""
    stored_block = int.from_bytes(raw[pos + 8:pos + 12], "little")
    stored_header = int.from_bytes(raw[pos + 12:pos + 16], "little")
    if crc32c_raw(raw[pos:pos + 12]) != stored_header:
        raise ValueError("block at %d: header checksum does not match" % pos)
    if crc32c_raw(raw[pos + 16:pos + overall]) != stored_block:
        raise ValueError("block at %d: block checksum does not match" % pos)


def walk_blocks_ex(raw, checksum, verify=False):
    """every storage block of a column file: dict(hoff, hlen, off, dlen, clen, rows, first, kind); clen > 0 = the content
    is stored bulk-compressed in clen bytes (AppendOnlyStorageFormat_GetSmallContentHeaderInfo /
    _GetBulkDenseContentHeaderInfo, cdbappendonlystorageformat.c:1327-1400, 1480-1560)"""
    out = []
    pos = 0
    n = len(raw)
    while pos < n:
        w0 = int.from_bytes(raw[pos:pos + 4], "little")
        w1 = int.from_bytes(raw[pos + 4:pos + 8], "little")
        kind = (w0 & 0x70000000) >> 28
        has_first = (w0 & 0x08000000) >> 27
        hlen = 8 + (8 if checksum else 0)
        clen = 0
        if kind in (1, 4):      # SmallContent / BulkDenseContent: same length fields
            rows = (w0 & 0x00FFFC00) >> 10
            dlen = ((w0 & 0x3FF) << 11) | ((w1 & 0xFFE00000) >> 21)
            clen = w1 & 0x001FFFFF
            if kind == 4:       # extension header behind the checksums: 30-bit row count
                rows = int.from_bytes(raw[pos + hlen + 4:pos + hlen + 8], "little") & 0x3FFFFFFF
                hlen += 8
        elif kind == 3:         # NonBulkDenseContent: 30-bit row count
            dlen = w0 & 0x001FFFFF
            rows = w1 & 0x3FFFFFFF
        else:
            raise ValueError("block at %d: header kind %d is not SmallContent / NonBulkDenseContent / BulkDenseContent" % (pos, kind))
        first = -1
        if has_first:
            first = int.from_bytes(raw[pos + hlen:pos + hlen + 8], "little", signed=True)
            hlen += 8
        stored = clen if clen else dlen
        if verify and checksum:
            verify_block_checksums(raw, pos, hlen + (stored + 7) // 8 * 8)
        out.append(dict(hoff=pos, hlen=hlen, off=pos + hlen, dlen=dlen, clen=clen, rows=rows, first=first, kind=kind))
        pos += hlen + (stored + 7) // 8 * 8
    return out


def walk_blocks(raw, checksum, verify=False):
    """[(content offset, content length, row count, first row number)] of an uncompressed column file"""
    out = []
    for b in walk_blocks_ex(raw, checksum, verify):
        if b["clen"]:
            raise ValueError("compressed block")
        out.append((b["off"], b["dlen"], b["rows"], b["first"]))
    return out


def block_contents(raw, checksum, verify=False, compresstype="zlib"):
    """[(datum stream block bytes, row count)]: AppendOnlyStorageRead_Content (cdbappendonlystorageread.c:1136-1320):
    the stored bytes, or for a bulk-compressed block what zlib's uncompress() makes of them (zlib_decompress,
    catalog/pg_compression.c:321-370; gp_decompress checks the length, storage/file/gp_compress.c:52-90)"""
    import zlib
    out = []
    for b in walk_blocks_ex(raw, checksum, verify):
        if b["clen"] and compresstype == "zstd":
            # zstd_decompress (gpcontrib/zstd/zstd_compression.c:142-175): ZSTD_decompressDCtx into dst_sz bytes
            import pyarrow as pa
            blk = pa.Codec("zstd").decompress(bytes(raw[b["off"]:b["off"] + b["clen"]]), decompressed_size=b["dlen"], asbytes=True)
        elif b["clen"]:
            blk = zlib.decompress(bytes(raw[b["off"]:b["off"] + b["clen"]]))
        if b["clen"]:
            if len(blk) != b["dlen"]:
                raise ValueError("block at %d: inflated to %d bytes, header says %d" % (b["hoff"], len(blk), b["dlen"]))
        else:
            blk = raw[b["off"]:b["off"] + b["dlen"]]
        out.append((blk, b["rows"]))
    return out


def _varint(buf, p):
    """DatumStreamInt32Compress_Decode: top two bits of the first byte = length - 1, big endian"""
    n = (buf[p] >> 6) + 1
    v = buf[p] & 0x3F
    for i in range(1, n):
        v = (v << 8) | buf[p + i]
    return v, n


class _Datums:
    """the datum area of one block, read front to back (AdvanceOrig / AdvanceDense pointer rules)"""

    def __init__(self, blk, p, end, typname, dscale):
        self.blk, self.p, self.end, self.typname, self.dscale = blk, p, end, typname, dscale
        self.typid, self.attlen, _, align, _ = TYPEINFO[typname]
        self.alignto = {"c": 1, "s": 2, "i": 4, "d": 8}[align]

    def next(self):
        blk, p = self.blk, self.p
        if self.attlen > 0:
            v = int.from_bytes(blk[p:p + self.attlen], "little", signed=self.typname != "bool")
            self.p = p + self.attlen
            return v
        b0 = blk[p]
        if b0 & 1:
            size = b0 >> 1
            body = blk[p + 1:p + size]
        else:
            size = (int.from_bytes(blk[p:p + 4], "little") >> 2) & 0x3FFFFFFF
            body = blk[p + 4:p + size]
        if self.typname in STRING_TYPES:
            v = bytes(body)
        else:
            v = numeric_from_bytes(body, self.dscale) if self.typname == "numeric" else (body[0] if len(body) else 32)
        p += size
        if p < self.end and blk[p] == 0:
            p = (p + self.alignto - 1) // self.alignto * self.alignto
        self.p = p
        return v


def decode_column(raw, typname, checksum, dscale=0, compresstype="zlib"):
    """(values, nulls): numeric -> scaled int64, bpchar -> first byte, fixed width -> the value"""
    vals, nulls = [], []
    for blk, rows in block_contents(raw, checksum, compresstype=compresstype):
        version, flags = int.from_bytes(blk[0:2], "little", signed=True), int.from_bytes(blk[2:4], "little")
        if version == 0:
            # DatumStreamBlock_Orig
            ndatum = int.from_bytes(blk[4:6], "little", signed=True)
            nullsz = int.from_bytes(blk[8:12], "little")
            sz = int.from_bytes(blk[12:16], "little")
            if ndatum != rows:
                raise ValueError("row counts disagree")
            p = 16
            bitmap = None
            if flags & 1:
                bitmap = blk[p:p + nullsz]
                p += nullsz
            p = (p + 7) // 8 * 8
            d = _Datums(blk, p, p + sz, typname, dscale)
            for r in range(rows):
                if bitmap is not None and (bitmap[r >> 3] >> (r & 7)) & 1:
                    vals.append(0)
                    nulls.append(1)
                else:
                    vals.append(d.next())
                    nulls.append(0)
            if d.p > d.end + d.alignto:
                raise ValueError("datum area overrun")
            continue
        if version not in (1, 2):
            raise ValueError("unknown datum stream block version %d" % version)
        # DatumStreamBlock_Dense (+ Rle_Extension)
        logical = int.from_bytes(blk[4:8], "little", signed=True)
        psize = int.from_bytes(blk[12:16], "little", signed=True)
        if logical != rows:
            raise ValueError("row counts disagree")
        p = 16
        rle = bool(flags & 2)
        delta = bool(flags & 4)
        if rle:
            nn_count, c_count, rc_count, rc_size = [int.from_bytes(blk[p + 4 * i:p + 4 * i + 4], "little") for i in range(4)]
            p += 16
        if delta:
            d_count, d_items, d_size = [int.from_bytes(blk[p + 4 * i:p + 4 * i + 4], "little") for i in range(3)]
            p += 12
        bitmap = None
        if flags & 1:
            nb = nn_count if rle else logical
            bitmap = blk[p:p + (nb + 7) // 8]
            p += (nb + 7) // 8
        if rle:
            cbitmap = blk[p:p + (c_count + 7) // 8]
            p += (c_count + 7) // 8
            rc = blk[p:p + rc_size]
            p += rc_size
        if delta:
            dbitmap = blk[p:p + (d_count + 7) // 8]
            p += (d_count + 7) // 8
            deltas = blk[p:p + d_size]
            p += d_size
        p = (p + 7) // 8 * 8
        d = _Datums(blk, p, p + psize, typname, dscale)
        produced = pos = di = rcp = dp = 0
        running = 0
        width = TYPEINFO[typname][1]
        while produced < logical:
            if bitmap is not None:
                isnull = (bitmap[pos >> 3] >> (pos & 7)) & 1
                pos += 1
                if isnull:
                    vals.append(0)
                    nulls.append(1)
                    produced += 1
                    continue
            rep = 0
            if rle and (cbitmap[di >> 3] >> (di & 7)) & 1:
                rep, n = _varint(rc, rcp)
                rcp += n
            if delta and (dbitmap[di >> 3] >> (di & 7)) & 1:
                # DatumStreamInt32CompressReserved3_Decode: 2 length bits, 1 "positive" bit, 29 value bits, big endian
                nb = (deltas[dp] >> 6) + 1
                mag = deltas[dp] & 0x1F
                for i in range(1, nb):
                    mag = (mag << 8) | deltas[dp + i]
                running = running + mag if deltas[dp] & 0x20 else running - mag
                dp += nb
                bits = 8 * width
                running &= (1 << bits) - 1
                v = running - (1 << bits) if running >> (bits - 1) else running
            else:
                v = d.next()
                running = v & ((1 << (8 * width)) - 1) if width > 0 else 0
            di += 1
            vals.extend([v] * (1 + rep))
            nulls.extend([0] * (1 + rep))
            produced += 1 + rep
        if produced != logical:
            raise ValueError("repeat counts overrun the block")
    if typname == "float8":
        return np.array(vals, dtype=np.int64).view(np.float64), np.array(nulls, dtype=np.uint8)
    if typname in STRING_TYPES:
        return [b"" if v == 0 else v for v in vals], np.array(nulls, dtype=np.uint8)
    return np.array(vals, dtype=np.int64), np.array(nulls, dtype=np.uint8)


# ------------------------------------------------------------------------------------------------
# visibility map entries (pg_aovisimap_<oid>: segno, first_row_no, visimap)
# ------------------------------------------------------------------------------------------------
VISIMAP_RANGE = 32768           # APPENDONLY_VISIMAP_MAX_RANGE (access/appendonly_visimap.h:36): rows per entry


def ref_visimap_entry(offsets, raw=False):
    """payload of pg_aovisimap.visimap (after the varlena length word) hiding the given row offsets of one entry,
    through the reference's Bitmap_Compress"""
    L = ref_lib()
    offs = np.ascontiguousarray(offsets, dtype=np.int32)
    out = (C.c_ubyte * 4200)()
    n = L.ref_visimap_entry_write(offs.ctypes.data, len(offs), 1 if raw else 0, out, 4200)
    if n < 0:
        raise RuntimeError("reference visimap writer failed")
    return bytes(out[:n])


def visimap_entry_blocks(payload):
    """the 32-bit bitmap blocks of one entry (bit j of block i = row offset 32 i + j is HIDDEN): int32 version 1, then
    BitmapDecompress_Init / _Decompress (utils/misc/bitmap_compression.c:31-52, 96-190) over an MSB-first bit stream
    (utils/misc/bitstream.c:52-70, 160-178): 1 bit compression type, 3 unused, 12 bits block count; type 0 = raw
    little-endian words from byte 2; type 1 = per block a 2-bit flag: 00 zero, 01 all ones, 11 raw 32 bits, 10 repeat
    the last block (8-bit count + 1 times in total)"""
    if len(payload) < 6 or int.from_bytes(payload[0:4], "little") != 1:
        raise ValueError("visimap entry: version")
    data = payload[4:]
    pos = 0

    def get(n):
        nonlocal pos
        v = 0
        for _ in range(n):
            if pos >> 3 >= len(data):
                raise ValueError("visimap entry: bit stream ends early")
            v = (v << 1) | ((data[pos >> 3] >> (7 - (pos & 7))) & 1)
            pos += 1
        return v
    ctype = get(1)
    get(3)
    count = get(12)
    if count > 1024:
        raise ValueError("visimap entry: block count")
    out = np.zeros(count, dtype=np.uint32)
    if ctype == 0:
        if 2 + 4 * count > len(data):
            raise ValueError("visimap entry: short raw bitmap")
        return np.frombuffer(data[2:2 + 4 * count], dtype="<u4").copy()
    last = 0
    repeat = 0
    for i in range(count):
        if repeat:
            repeat -= 1
        else:
            flag = get(2)
            if flag == 0:
                last = 0
            elif flag == 1:
                last = 0xFFFFFFFF
            elif flag == 3:
                last = get(32)
            else:
                if i == 0:
                    raise ValueError("visimap entry: repeat before any block")
                repeat = get(8)
        out[i] = last
    if repeat:
        raise ValueError("visimap entry: repeat runs past the last block")
    return out


def visimap_visible(raw, checksum, entries):
    """one bool per row of a column file, in file order: AppendOnlyVisimap_IsVisible
    (access/appendonly/appendonly_visimap.c:198 -> AppendOnlyVisimapEntry_IsVisible, appendonly_visimap_entry.c:451-493)
    for row number = the block's firstRowNum + position in the block.  entries: {first_row_no: payload or None}; a row
    whose range has no entry, or a NULL visimap, is visible."""
    decoded = {}
    out = []
    for b in walk_blocks_ex(raw, checksum):
        if b["first"] < 0:
            raise ValueError("block without a first row number")
        for i in range(b["rows"]):
            rn = b["first"] + i
            first = rn // VISIMAP_RANGE * VISIMAP_RANGE
            if first not in entries or entries[first] is None:
                out.append(True)
                continue
            if first not in decoded:
                decoded[first] = visimap_entry_blocks(entries[first])
            blocks = decoded[first]
            off = rn - first
            out.append(not (off // 32 < len(blocks) and (int(blocks[off // 32]) >> (off % 32)) & 1))
    return np.array(out, dtype=bool)
