/*
 * oracle/ref_aocs.c - drives the REFERENCE's own AOCS block writer to produce golden column files.
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile (only where /root/reference exists) into
 * oracle/_ref/libaocs_ref.so together with three reference sources compiled where they lie:
 *     src/backend/utils/datumstream/datumstreamblock.c    datum stream blocks (Orig / Dense / RLE / delta)
 *     src/backend/cdb/cdbappendonlystorageformat.c        Append-Only storage block headers + checksums
 *     src/port/pg_crc32c_sb8.c                            CRC-32C
 * with the stand-in generated headers under oracle/ref_shim/ (the reference's configure / genbki
 * are not run).  No reference source is copied: this file only
 *   - stubs the backend services those sources call (palloc, ereport, a few GUC flags, encryption
 *     hooks) so that they link outside a backend, and
 *   - replays the put / flush loop of aocs_insert_values (access/aocs/aocsam.c:1583-1640) ->
 *     datumstreamwrite_put / datumstreamwrite_block_orig (utils/datumstream/datumstream.c:300,880-923)
 *     -> AppendOnlyStorageWrite_FinishBuffer (cdb/cdbappendonlystoragewrite.c:1183-1300, the
 *     uncompressed small-content branch) for one column,
 * so the bytes in the golden fixtures are what the reference writes into a column's segment file.
 * tests/golden/make_aocs_golden.py calls it; the fixtures it writes are what travels.
 */
#include "postgres.h"

#include <setjmp.h>
#include <stdarg.h>
#include <zlib.h>

#include "catalog/pg_appendonly.h"
#include "cdb/cdbappendonlystorage.h"
#include "cdb/cdbappendonlystorageformat.h"
#include "utils/datumstreamblock.h"
#include "utils/bitmap_compression.h"
#include "utils/numeric.h"

/* port.h routes the printf family to the reference's own src/port implementations, which are not linked here */
#undef vsnprintf
#undef snprintf
#undef vsprintf
#undef sprintf
#undef printf
#undef fprintf
#undef vfprintf

/* ---- backend services the three sources reach for ---- */
MemoryContext CurrentMemoryContext = NULL;
bool		Debug_appendonly_print_insert = false;
bool		Debug_appendonly_print_insert_tuple = false;
bool		Debug_appendonly_print_scan = false;
bool		Debug_appendonly_print_scan_tuple = false;
bool		Debug_appendonly_print_storage_headers = false;
bool		Debug_appendonly_print_verify_write_block = false;
bool		Debug_datumstream_block_read_check_integrity = true;
bool		Debug_datumstream_block_write_check_integrity = true;
bool		Debug_datumstream_write_print_small_varlena_info = false;
bool		Debug_datumstream_write_use_small_initial_buffers = false;
bool		Debug_datumstream_read_check_large_varlena_integrity = false;
bool		Debug_datumstream_read_print_varlena_info = false;
bool		FileEncryptionEnabled = false;

jmp_buf		ref_jmp;			/* shared with ref_tupser.c */
static char ref_errbuf[512];
static int	ref_elevel;

/* a backend function the compiled reference files link against but these drivers never reach */
void
ref_abort(const char *what)
{
	snprintf(ref_errbuf, sizeof(ref_errbuf), "oracle/_ref: %s is a stub", what);
	longjmp(ref_jmp, 1);
}

void	   *palloc(Size size) { return malloc(size ? size : 1); }
void	   *palloc0(Size size) { return calloc(1, size ? size : 1); }
void	   *repalloc(void *p, Size size) { return realloc(p, size ? size : 1); }
void		pfree(void *p) { free(p); }

bool
errstart(int elevel, const char *domain)
{
	(void) domain;
	ref_elevel = elevel;
	return elevel >= ERROR;		/* LOG / DEBUG chatter is dropped */
}

bool
errstart_cold(int elevel, const char *domain)
{
	return errstart(elevel, domain);
}

void
errfinish(const char *filename, int lineno, const char *funcname)
{
	(void) funcname;
	if (ref_elevel >= ERROR)
	{
		size_t		n = strlen(ref_errbuf);

		snprintf(ref_errbuf + n, sizeof(ref_errbuf) - n, " (%s:%d)", filename, lineno);
		longjmp(ref_jmp, 1);
	}
}

void
errmsg(const char *fmt,...)
{
	va_list		ap;

	va_start(ap, fmt);
	vsnprintf(ref_errbuf, sizeof(ref_errbuf), fmt, ap);
	va_end(ap);
}

void
errmsg_internal(const char *fmt,...)
{
	va_list		ap;

	va_start(ap, fmt);
	vsnprintf(ref_errbuf, sizeof(ref_errbuf), fmt, ap);
	va_end(ap);
}

void		errdetail(const char *fmt,...) { (void) fmt; }
void		errdetail_internal(const char *fmt,...) { (void) fmt; }
void		errcode(int sqlerrcode) { (void) sqlerrcode; }
int			errprintstack(bool printstack) { (void) printstack; return 0; }

int
pg_sprintf(char *str, const char *fmt,...)
{
	va_list		ap;
	int			n;

	va_start(ap, fmt);
	n = vsprintf(str, fmt, ap);
	va_end(ap);
	return n;
}

int
pg_snprintf(char *str, size_t count, const char *fmt,...)
{
	va_list		ap;
	int			n;

	va_start(ap, fmt);
	n = vsnprintf(str, count, fmt, ap);
	va_end(ap);
	return n;
}

char *
psprintf(const char *fmt,...)
{
	char	   *buf = malloc(1024);
	va_list		ap;

	va_start(ap, fmt);
	vsnprintf(buf, 1024, fmt, ap);
	va_end(ap);
	return buf;
}

void		EncryptAOBLock(unsigned char *data_buf, const int buf_len, RelFileNode *file_node) { (void) data_buf; (void) buf_len; (void) file_node; }
void		DecryptAOBlock(unsigned char *data_buf, const int buf_len, RelFileNode *file_node) { (void) data_buf; (void) buf_len; (void) file_node; }

/* access/common/detoast.c:652 for the plain (never toasted, never compressed) datums the fixtures hold */
void
varattrib_untoast_ptr_len(Datum d, char **datastart, int *len, void **tofree)
{
	struct varlena *va = (struct varlena *) DatumGetPointer(d);

	*tofree = NULL;
	if (VARATT_IS_SHORT(va))
	{
		*len = VARSIZE_SHORT(va) - VARHDRSZ_SHORT;
		*datastart = VARDATA_SHORT(va);
	}
	else
	{
		*len = VARSIZE(va) - VARHDRSZ;
		*datastart = VARDATA(va);
	}
}

const char *
ref_aocs_last_error(void)
{
	return ref_errbuf;
}

/* ---- numeric varlena datums (by reference) through the reference's own header macros ---- */

/* read back a numeric datum with utils/numeric.h's accessors: sign, dscale, weight, digits */
int
ref_numeric_inspect(const unsigned char *varlena4, int *sign, int *dscale, int *weight, int *ndigits, int16 *digits, int maxdigits)
{
	Numeric		num = (Numeric) varlena4;
	int			n = NUMERIC_NDIGITS(num);

	if (NUMERIC_IS_SPECIAL(num))
		return -1;
	*sign = NUMERIC_SIGN(num) == NUMERIC_NEG;
	*dscale = NUMERIC_DSCALE(num);
	*weight = NUMERIC_WEIGHT(num);
	*ndigits = n;
	for (int i = 0; i < n && i < maxdigits; i++)
		digits[i] = NUMERIC_DIGITS(num)[i];
	return NUMERIC_HEADER_IS_SHORT(num) ? 1 : 0;
}

int64		ref_aocs_write_column_ex(int typid, int attlen, int byval, int align, int storage, int checksum, int blocksize, int rle,
									 const int64 *values, const unsigned char *varbuf, const unsigned char *nulls, int64 n,
									 unsigned char *out, int64 outcap, int64 *nblocks_out);
int64		ref_aocs_write_column_z(int typid, int attlen, int byval, int align, int storage, int checksum, int blocksize, int rle,
									int zlevel, const int64 *values, const unsigned char *varbuf, const unsigned char *nulls, int64 n,
									unsigned char *out, int64 outcap, int64 *nblocks_out);

/*
 * Write one column.  `values`: by-value datums (attlen 1/2/4/8), or for attlen -1 offsets into
 * `varbuf` of 4-byte-header varlenas.  Returns bytes written into out (the column's segment-file
 * content: storage blocks back to back), -1 on error (ref_aocs_last_error).
 */
int64
ref_aocs_write_column(int typid, int attlen, int byval, int align, int storage, int checksum, int blocksize,
					  const int64 *values, const unsigned char *varbuf, const unsigned char *nulls, int64 n,
					  unsigned char *out, int64 outcap, int64 *nblocks_out)
{
	return ref_aocs_write_column_ex(typid, attlen, byval, align, storage, checksum, blocksize, 0, values, varbuf, nulls, n, out, outcap,
									nblocks_out);
}

/* rle != 0: compresstype = rle_type, compresslevel 1 (init_datumstream_info, datumstream.c:396-419):
 * DatumStreamVersion_Dense_Enhanced blocks with RLE done by the datum stream layer, no bulk compression;
 * blocks of more than 16383 rows get the NonBulkDenseContent storage header (datumstreamwrite_block_dense,
 * datumstream.c:925-978).  rle == 2 adds delta range encoding, which init_datumstream_info turns on for
 * int4 / int8 / date / time / timestamp columns (is_deltarange_compression_supported, datumstream.c:330-360). */
int64
ref_aocs_write_column_ex(int typid, int attlen, int byval, int align, int storage, int checksum, int blocksize, int rle,
						 const int64 *values, const unsigned char *varbuf, const unsigned char *nulls, int64 n,
						 unsigned char *out, int64 outcap, int64 *nblocks_out)
{
	return ref_aocs_write_column_z(typid, attlen, byval, align, storage, checksum, blocksize, rle, 0, values, varbuf, nulls, n, out,
								   outcap, nblocks_out);
}

/* zlevel > 0: bulk compression by the storage layer, restating AppendOnlyStorageWrite_CompressAppend
 * (cdbappendonlystoragewrite.c:1015-1160; that file needs the backend's file layer and is not compiled here):
 * the content goes through zlib's compress2() exactly as zlib_compress does (catalog/pg_compression.c:272-318, with
 * Z_BUF_ERROR meaning "did not fit: store as is"), is kept compressed only when shorter than the source, and the
 * REFERENCE's own header makers record the compressed length.  Dense blocks beyond 16383 rows take the
 * BulkDenseContent header when bulk compression is on (datumstreamwrite_block_dense, datumstream.c:944-956).
 * compresstype=zlib, compresslevel=L: rle 0, zlevel L.  rle_type compresslevel 2 / 3 / 4: rle 1|2, zlevel 1 / 5 / 9
 * (init_datumstream_info, datumstream.c:412-436). */
typedef int (*ref_compress_cb) (const unsigned char *src, int srclen, unsigned char *dst, int dstcap);

static int64 ref_write_impl(int typid, int attlen, int byval, int align, int storage, int checksum, int blocksize, int rle, int zlevel,
							ref_compress_cb cb, const int64 *values, const unsigned char *varbuf, const unsigned char *nulls, int64 n,
							unsigned char *out, int64 outcap, int64 *nblocks_out);

int64
ref_aocs_write_column_z(int typid, int attlen, int byval, int align, int storage, int checksum, int blocksize, int rle, int zlevel,
						const int64 *values, const unsigned char *varbuf, const unsigned char *nulls, int64 n,
						unsigned char *out, int64 outcap, int64 *nblocks_out)
{
	return ref_write_impl(typid, attlen, byval, align, storage, checksum, blocksize, rle, zlevel, NULL, values, varbuf, nulls, n, out,
						  outcap, nblocks_out);
}

/* Bulk compression through a caller-supplied compressor, for compression libraries this container cannot link (zstd:
 * the reference's zstd_compress, gpcontrib/zstd/zstd_compression.c:104-140, calls ZSTD_compressCCtx and reports
 * "did not fit" as dst_used = src_sz).  cb returns the compressed length, or srclen when dstcap was too small. */
int64
ref_aocs_write_column_cb(int typid, int attlen, int byval, int align, int storage, int checksum, int blocksize, int rle,
						 ref_compress_cb cb, const int64 *values, const unsigned char *varbuf, const unsigned char *nulls, int64 n,
						 unsigned char *out, int64 outcap, int64 *nblocks_out)
{
	return ref_write_impl(typid, attlen, byval, align, storage, checksum, blocksize, rle, 1, cb, values, varbuf, nulls, n, out, outcap,
						  nblocks_out);
}

static int64
ref_write_impl(int typid, int attlen, int byval, int align, int storage, int checksum, int blocksize, int rle, int zlevel,
			   ref_compress_cb cb, const int64 *values, const unsigned char *varbuf, const unsigned char *nulls, int64 n,
			   unsigned char *out, int64 outcap, int64 *nblocks_out)
{
	DatumStreamBlockWrite dsw;
	DatumStreamTypeInfo ti;
	RelFileNode node;
	const int	version = AOSegfileFormatVersion_GetLatest();
	const int	hdrlen = AppendOnlyStorageFormat_RegularHeaderLenNeeded(checksum != 0) + (int) sizeof(int64);	/* + firstRowNum */
	int64		pos = 0;
	int64		first_row = 1;
	int64		nblocks = 0;
	unsigned char *blockbuf = malloc((size_t) blocksize + 64);
	unsigned char *contentbuf = malloc((size_t) blocksize + 64);

	memset(&node, 0, sizeof(node));
	memset(&dsw, 0, sizeof(dsw));
	ti.datumlen = attlen;
	ti.typid = typid;
	ti.typstorage = (char) storage;
	ti.align = (char) align;
	ti.byval = byval != 0;
	ref_errbuf[0] = 0;
	if (setjmp(ref_jmp))
	{
		free(blockbuf);
		return -1;
	}
	/* create_datumstreamwrite (datumstream.c:588-632) */
	if (rle)
		DatumStreamBlockWrite_Init(&dsw, &ti, DatumStreamVersion_Dense_Enhanced, true, rle >= 2,
								   AOSmallContentHeader_MaxRowCount, AONonBulkDenseContentHeader_MaxLargeRowCount,
								   blocksize - (AoHeader_LongSize + (checksum ? 8 : 0) + 8),
								   NULL, NULL, NULL, NULL, &node);
	else
		DatumStreamBlockWrite_Init(&dsw, &ti, DatumStreamVersion_Original, false, false,
								   AOSmallContentHeader_MaxRowCount, AOSmallContentHeader_MaxRowCount, blocksize - hdrlen,
								   NULL, NULL, NULL, NULL, &node);

#define FLUSH() \
	do { \
		int			rowCount = DatumStreamBlockWrite_Nth(&dsw); \
		int64		contentLen; \
		int32		rounded; \
		int32		compressedLen = 0; \
		if (rowCount > 0) \
		{ \
			const int	bulk = rowCount > AOSmallContentHeader_MaxRowCount && zlevel > 0; \
			const int	hl = hdrlen + (bulk ? AoHeader_RegularSize : 0);	/* + the extension header */ \
			memset(blockbuf, 0, (size_t) blocksize + 64); \
			contentLen = DatumStreamBlockWrite_Block(&dsw, contentbuf, &node); \
			if (cb) \
				compressedLen = cb(contentbuf, (int) contentLen, blockbuf + hl, blocksize - hl); \
			else if (zlevel > 0) \
			{ \
				unsigned long used = (unsigned long) (blocksize - hl); \
				int			zrc = compress2(blockbuf + hl, &used, contentbuf, (unsigned long) contentLen, zlevel); \
				compressedLen = zrc == Z_OK ? (int32) used : (int32) contentLen; \
				if (zrc != Z_OK && zrc != Z_BUF_ERROR) \
				{ \
					snprintf(ref_errbuf, sizeof(ref_errbuf), "compress2 failed: %d", zrc); \
					free(blockbuf); free(contentbuf); \
					return -1; \
				} \
			} \
			if (zlevel <= 0 || compressedLen >= contentLen) \
			{ \
				memset(blockbuf + hl, 0, (size_t) blocksize + 64 - hl); \
				memcpy(blockbuf + hl, contentbuf, (size_t) contentLen); \
				compressedLen = 0; \
			} \
			rounded = AOStorage_RoundUp(compressedLen ? compressedLen : (int32) contentLen, version); \
			memset(blockbuf + hl + (compressedLen ? compressedLen : (int32) contentLen), 0, \
				   (size_t) (rounded - (compressedLen ? compressedLen : (int32) contentLen))); \
			if (rowCount <= AOSmallContentHeader_MaxRowCount) \
				AppendOnlyStorageFormat_MakeSmallContentHeader(blockbuf, checksum != 0, true, version, first_row, 1 /* AOCSBK_BLOCK */, \
															   rowCount, (int32) contentLen, compressedLen); \
			else if (bulk) \
				AppendOnlyStorageFormat_MakeBulkDenseContentHeader(blockbuf, checksum != 0, true, version, first_row, 1, \
																   rowCount, (int32) contentLen, compressedLen); \
			else \
				AppendOnlyStorageFormat_MakeNonBulkDenseContentHeader(blockbuf, checksum != 0, true, version, first_row, 1, \
																	  rowCount, (int32) contentLen); \
			if (pos + hl + rounded > outcap) \
			{ \
				snprintf(ref_errbuf, sizeof(ref_errbuf), "output buffer too small"); \
				free(blockbuf); free(contentbuf); \
				return -1; \
			} \
			memcpy(out + pos, blockbuf, (size_t) hl + (size_t) rounded); \
			pos += hl + rounded; \
			first_row += rowCount; \
			nblocks++; \
			DatumStreamBlockWrite_GetReady(&dsw); \
		} \
	} while (0)

	for (int64 i = 0; i < n; i++)
	{
		bool		isnull = nulls && nulls[i];
		Datum		d = 0;
		void	   *toFree = NULL;
		int			err;

		if (!isnull)
			d = attlen == -1 ? PointerGetDatum(varbuf + values[i]) : (Datum) values[i];
		err = DatumStreamBlockWrite_Put(&dsw, d, isnull, &toFree);
		if (err < 0)
		{
			FLUSH();
			err = DatumStreamBlockWrite_Put(&dsw, d, isnull, &toFree);
			if (err < 0)
			{
				snprintf(ref_errbuf, sizeof(ref_errbuf), "datum %lld does not fit an empty block", (long long) i);
				free(blockbuf);
				return -1;
			}
		}
	}
	FLUSH();
	DatumStreamBlockWrite_Finish(&dsw);
	free(blockbuf);
	free(contentbuf);
	if (nblocks_out)
		*nblocks_out = nblocks;
	return pos;
}

/* compressed length of the block ref_aocs_block_info looked at last (0 = stored as is) */
static int32 ref_last_compressed_len;

int
ref_aocs_last_compressed_len(void)
{
	return ref_last_compressed_len;
}

/* parse one storage block header with the reference's own accessors (for cross-checking a walker) */
int
ref_aocs_block_info(const unsigned char *hdr, int checksum, int *header_len, int *row_count, int *data_len, int64 *first_row,
					int *header_kind)
{
	const int	version = AOSegfileFormatVersion_GetLatest();
	AOHeaderCheckError e;
	int32		overall = 0,
				offset = 0,
				uncompressed = 0,
				compressed = 0;
	int			exec_kind = 0;
	bool		has_first = false;
	int64		fr = -1;
	int			rc = 0;
	int32		hlen = 0;

	if (setjmp(ref_jmp))
		return -1;
	e = AppendOnlyStorageFormat_GetHeaderInfo((uint8 *) hdr, checksum != 0, header_kind, &hlen);
	if (e != AOHeaderCheckOk)
		return -2;
	{
		bool		is_compressed = false;

		if (*header_kind == AoHeaderKind_NonBulkDenseContent)
			e = AppendOnlyStorageFormat_GetNonBulkDenseContentHeaderInfo((uint8 *) hdr, hlen, checksum != 0, 1 << 21, &overall, &offset,
																		 &uncompressed, &exec_kind, &has_first, version, &fr, &rc);
		else if (*header_kind == AoHeaderKind_BulkDenseContent)
			e = AppendOnlyStorageFormat_GetBulkDenseContentHeaderInfo((uint8 *) hdr, hlen, checksum != 0, 1 << 21, &overall, &offset,
																	  &uncompressed, &exec_kind, &has_first, version, &fr, &rc,
																	  &is_compressed, &compressed);
		else
			e = AppendOnlyStorageFormat_GetSmallContentHeaderInfo((uint8 *) hdr, hlen, checksum != 0, 1 << 21, &overall, &offset,
																  &uncompressed, &exec_kind, &has_first, version, &fr, &rc,
																  &is_compressed, &compressed);
	}
	if (e != AOHeaderCheckOk)
		return -3;
	ref_last_compressed_len = compressed;
	*header_len = offset;
	*row_count = rc;
	*data_len = uncompressed;
	*first_row = fr;
	return 0;
}

/* verify the checksums of one block with the reference's routines: 0 = both good */
int
ref_aocs_verify_block(const unsigned char *hdr, int overall_len)
{
	pg_crc32	stored = 0,
				computed = 0;

	if (setjmp(ref_jmp))
		return -1;
	if (!AppendOnlyStorageFormat_VerifyHeaderChecksum((uint8 *) hdr, &stored, &computed))
		return 1;
	if (!AppendOnlyStorageFormat_VerifyBlockChecksum((uint8 *) hdr, overall_len, &stored, &computed))
		return 2;
	return 0;
}


/* ---- visibility map entries: the reference's own bitmap codec (utils/misc/bitmap_compression.c, bitstream.c) ---- */

/*
 * pg_aovisimap.visimap payload (after the varlena length word) for a set of hidden row offsets within one entry's
 * 32768-row range, as AppendOnlyVisimapEntry_WriteData builds it (appendonly_visimap_entry.c:282-316): int32 version 1,
 * then Bitmap_Compress(DEFAULT) over the bitmapset's words as 32-bit blocks.  The bitmapset machinery (nodes/bitmapset.c)
 * is not compiled here; its word count is restated: grown in powers of two to cover the highest member
 * (AppendOnlyVisimapEntry_HideTuple :556-566), one block when only the low 32 bits of a single word are used, else two
 * blocks per 64-bit word (BitmapCompress_CalculateBlockCounts, bitmap_compression.c:415-460).
 * Returns bytes written, -1 on error.
 */
int
ref_visimap_entry_write(const int *offsets, int noffsets, int raw, unsigned char *out, int outcap)
{
	uint32		blocks[1024];
	int			maxoff = -1;
	int			nwords64 = 0;
	int			blockCount;
	int			n;

	if (setjmp(ref_jmp))
		return -1;
	memset(blocks, 0, sizeof(blocks));
	for (int i = 0; i < noffsets; i++)
	{
		if (offsets[i] < 0 || offsets[i] >= 32768)
			return -1;
		blocks[offsets[i] / 32] |= 1u << (offsets[i] % 32);
		if (offsets[i] > maxoff)
			maxoff = offsets[i];
	}
	if (maxoff >= 0)
	{
		nwords64 = 1;
		while (nwords64 * 64 <= maxoff)
			nwords64 *= 2;
	}
	blockCount = nwords64 == 0 ? 0 : (nwords64 == 1 && blocks[1] == 0) ? 1 : nwords64 * 2;
	if (outcap < 4 + 2 + 4 * blockCount + 8)
		return -1;
	memset(out, 0, (size_t) outcap);
	out[0] = 1;					/* version, little endian int32 */
	/* raw != 0: BITMAP_COMPRESSION_TYPE_NO, which the decompressor also accepts (bitmap_compression.c:118-123) */
	n = Bitmap_Compress(raw ? BITMAP_COMPRESSION_TYPE_NO : BITMAP_COMPRESSION_TYPE_DEFAULT, blocks, blockCount, out + 4, 2 + 4 * blockCount);
	return n < 0 ? -1 : 4 + n;
}

/* the reference's decompressor over such a payload: 32-bit blocks into out, returns the block count or -1 */
int
ref_visimap_entry_read(const unsigned char *payload, int len, uint32 *out, int outcap)
{
	BitmapDecompressState st;

	if (setjmp(ref_jmp))
		return -1;
	if (len < 4 + 2 || payload[0] != 1)
		return -1;
	if (!BitmapDecompress_Init(&st, (unsigned char *) payload + 4, len - 4) || BitmapDecompress_HasError(&st))
		return -1;
	if (BitmapDecompress_GetBlockCount(&st) > outcap)
		return -1;
	BitmapDecompress_Decompress(&st, out, outcap);
	return BitmapDecompress_GetBlockCount(&st);
}
