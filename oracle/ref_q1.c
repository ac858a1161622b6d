/*
 * oracle/ref_q1.c - TPC-H Q1's hot loop driven through the REFERENCE's own functions, as a CPU baseline and a second opinion.
 *
 * TEST / BENCH INFRASTRUCTURE ONLY (linked into oracle/_ref/libexec_ref.so, see oracle/Makefile).  The reference server cannot
 * be built here (no bison / flex), so its executor glue - ExecScan, ExecAgg, the expression interpreter, the tuple hash table -
 * is not available; everything those call PER ROW on this path is, compiled where it lies under /root/reference:
 *     scan        utils/datumstream/datumstreamblock.c   DatumStreamBlockRead_GetReady / _Advance / _Get over column files in the
 *                 cdb/cdbappendonlystorageformat.c        Append-Only storage format (headers, CRC-32C), written here by the reference's
 *                                                         own block writer - what aocs_getnext does per projected column (aocsam.c:1138)
 *     expressions utils/adt/numeric.c                     numeric_sub / numeric_add / numeric_mul through the fmgr call convention
 *     grouping    utils/adt/varchar.c, common/hashfn.c    hashbpchar per key, rotate-xor, murmurhash32 (execGrouping.c:473-495); bpchareq
 *     aggregates  utils/adt/numeric.c                     numeric_avg_accum (one shared state for sum(x) and avg(x) of the same x, as
 *                                                         nodeAgg.c's find_compatible_trans shares them), numeric_sum, numeric_avg
 * The glue restated here is the minimum around those calls: the block-to-block loop of datumstreamread_block
 * (datumstream.c:1364), the qual `l_shipdate <= date` as a DateADT comparison, a small open-addressing table in place of
 * simplehash, and a per-row reset of the allocation arena in place of ResetExprContext (execScan.c:195).  Every cost the
 * reference pays per row through these functions (varlena numerics, palloc per result, fmgr calls) is paid here too; costs of the
 * glue that is missing (slot handling, interpreter dispatch) are NOT - so this is a lower bound on the reference's CPU time.
 */
#include "postgres.h"

#include <setjmp.h>

#include "catalog/pg_collation.h"
#include "cdb/cdbappendonlystorage.h"
#include "cdb/cdbappendonlystorageformat.h"
#include "common/hashfn.h"
#include "fmgr.h"
#include "utils/datumstreamblock.h"
#include "utils/fmgrprotos.h"
#include "utils/numeric.h"

#undef snprintf
#undef vsnprintf
#undef sprintf
#undef printf
#undef fprintf

/* ---- backend globals / services datumstreamblock.c and cdbappendonlystorageformat.c reach for (as in ref_aocs.c) ---- */
bool		Debug_appendonly_print_insert = false;
bool		Debug_appendonly_print_insert_tuple = false;
bool		Debug_appendonly_print_scan = false;
bool		Debug_appendonly_print_scan_tuple = false;
bool		Debug_appendonly_print_storage_headers = false;
bool		Debug_appendonly_print_verify_write_block = false;
bool		Debug_datumstream_block_read_check_integrity = false;
bool		Debug_datumstream_block_write_check_integrity = false;
bool		Debug_datumstream_write_print_small_varlena_info = false;
bool		Debug_datumstream_write_use_small_initial_buffers = false;
bool		Debug_datumstream_read_check_large_varlena_integrity = false;
bool		Debug_datumstream_read_print_varlena_info = false;
bool		FileEncryptionEnabled = false;

void		EncryptAOBLock(unsigned char *data_buf, const int buf_len, RelFileNode *file_node) { (void) data_buf; (void) buf_len; (void) file_node; }
void		DecryptAOBlock(unsigned char *data_buf, const int buf_len, RelFileNode *file_node) { (void) data_buf; (void) buf_len; (void) file_node; }

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

extern void ref_exec_abort(const char *what);
extern void ref_arena_enable(int on);
extern void ref_arena_reset(void);
extern jmp_buf *ref_exec_jmp(void);

/* ---- a column file in memory ---- */
#define NCOLS 7
enum { C_QTY, C_EXT, C_DISC, C_TAX, C_SHIPDATE, C_RF, C_LS };

typedef struct ColFile
{
	unsigned char *bytes;
	int64		len;
	int64		cap;
	DatumStreamTypeInfo ti;
} ColFile;

typedef struct Q1Table
{
	int64		nrows;
	int			checksum;
	int			blocksize;
	ColFile		col[NCOLS];
} Q1Table;

typedef struct ColWriter
{
	ColFile    *f;
	DatumStreamBlockWrite dsw;
	RelFileNode node;
	int			checksum;
	int			blocksize;
	int			hdrlen;
	int64		first_row;
	unsigned char *content;
} ColWriter;

static void
file_append(ColFile *f, const unsigned char *p, int64 n)
{
	if (f->len + n > f->cap)
	{
		f->cap = (f->len + n) * 2 + 65536;
		f->bytes = realloc(f->bytes, (size_t) f->cap);
	}
	memcpy(f->bytes + f->len, p, (size_t) n);
	f->len += n;
}

/* AppendOnlyStorageWrite_FinishBuffer's uncompressed small-content branch (cdbappendonlystoragewrite.c:1183-1300) */
static void
writer_flush(ColWriter *w)
{
	const int	version = AOSegfileFormatVersion_GetLatest();
	int			rowCount = DatumStreamBlockWrite_Nth(&w->dsw);
	int64		contentLen;
	int32		rounded;
	unsigned char *block;

	if (rowCount <= 0)
		return;
	contentLen = DatumStreamBlockWrite_Block(&w->dsw, w->content, &w->node);
	rounded = AOStorage_RoundUp((int32) contentLen, version);
	block = calloc(1, (size_t) w->hdrlen + (size_t) rounded + 64);
	memcpy(block + w->hdrlen, w->content, (size_t) contentLen);
	AppendOnlyStorageFormat_MakeSmallContentHeader(block, w->checksum != 0, true, version, w->first_row, 1 /* AOCSBK_BLOCK */ ,
												   rowCount, (int32) contentLen, 0);
	file_append(w->f, block, w->hdrlen + rounded);
	free(block);
	w->first_row += rowCount;
	DatumStreamBlockWrite_GetReady(&w->dsw);
}

static void
writer_open(ColWriter *w, ColFile *f, int checksum, int blocksize)
{
	memset(w, 0, sizeof(*w));
	w->f = f;
	w->checksum = checksum;
	w->blocksize = blocksize;
	w->hdrlen = AppendOnlyStorageFormat_RegularHeaderLenNeeded(checksum != 0) + (int) sizeof(int64);
	w->first_row = 1;
	w->content = malloc((size_t) blocksize + 64);
	/* create_datumstreamwrite (datumstream.c:588-632), compresstype none */
	DatumStreamBlockWrite_Init(&w->dsw, &f->ti, DatumStreamVersion_Original, false, false,
							   AOSmallContentHeader_MaxRowCount, AOSmallContentHeader_MaxRowCount, blocksize - w->hdrlen,
							   NULL, NULL, NULL, NULL, &w->node);
}

static void
writer_put(ColWriter *w, Datum d)
{
	void	   *toFree = NULL;

	if (DatumStreamBlockWrite_Put(&w->dsw, d, false, &toFree) < 0)
	{
		writer_flush(w);
		if (DatumStreamBlockWrite_Put(&w->dsw, d, false, &toFree) < 0)
			ref_exec_abort("datum does not fit an empty block");
	}
}

static void
writer_close(ColWriter *w)
{
	writer_flush(w);
	DatumStreamBlockWrite_Finish(&w->dsw);
	free(w->content);
}

static void
set_typeinfo(DatumStreamTypeInfo *ti, Oid typid, int attlen, bool byval, char align, char storage)
{
	ti->datumlen = attlen;
	ti->typid = typid;
	ti->typstorage = storage;
	ti->align = align;
	ti->byval = byval;
}

/* ---- exported API ---- */
void	   *ref_q1_load(int64 n, const int64 *qty, const int64 *ext, const int64 *disc, const int64 *tax, const int32 *shipdate,
						const uint8 *rf, const uint8 *ls, int checksum, int blocksize);
int64		ref_q1_file_bytes(void *h);
int			ref_q1_run(void *h, int32 cutoff, char *out, int cap, int64 *rows_passed);
void		ref_q1_free(void *h);

/*
 * lineitem's seven Q1 columns as AOCS column files: numeric(15,2) for the four decimals (scaled x100 in; the datum is what
 * numeric_in would have produced, built by int64_div_fast_to_numeric), date, character(1) x 2.  Untimed set-up.
 */
void *
ref_q1_load(int64 n, const int64 *qty, const int64 *ext, const int64 *disc, const int64 *tax, const int32 *shipdate,
			const uint8 *rf, const uint8 *ls, int checksum, int blocksize)
{
	Q1Table    *t = calloc(1, sizeof(Q1Table));
	const int64 *dec[4] = {qty, ext, disc, tax};

	if (setjmp(*ref_exec_jmp()))
		return NULL;
	t->nrows = n;
	t->checksum = checksum;
	t->blocksize = blocksize;
	ref_arena_enable(0);		/* the writer keeps (and grows) palloc'd buffers across rows: plain malloc here */
	for (int c = 0; c < 4; c++)
	{
		ColWriter	w;

		set_typeinfo(&t->col[c].ti, 1700 /* NUMERICOID */ , -1, false, 'i', 'm');
		writer_open(&w, &t->col[c], checksum, blocksize);
		for (int64 i = 0; i < n; i++)
		{
			Numeric		num = int64_div_fast_to_numeric(dec[c][i], 2);

			writer_put(&w, NumericGetDatum(num));
			pfree(num);
		}
		writer_close(&w);
	}
	{
		ColWriter	w;

		set_typeinfo(&t->col[C_SHIPDATE].ti, 1082 /* DATEOID */ , 4, true, 'i', 'p');
		writer_open(&w, &t->col[C_SHIPDATE], checksum, blocksize);
		for (int64 i = 0; i < n; i++)
			writer_put(&w, Int32GetDatum(shipdate[i]));
		writer_close(&w);
	}
	for (int c = C_RF; c <= C_LS; c++)
	{
		ColWriter	w;
		const uint8 *src = c == C_RF ? rf : ls;
		struct
		{
			int32		hdr;
			char		data[4];
		}			v;

		set_typeinfo(&t->col[c].ti, 1042 /* BPCHAROID */ , -1, false, 'i', 'x');
		writer_open(&w, &t->col[c], checksum, blocksize);
		for (int64 i = 0; i < n; i++)
		{
			memset(&v, 0, sizeof(v));
			SET_VARSIZE(&v, VARHDRSZ + 1);
			v.data[0] = (char) src[i];
			writer_put(&w, PointerGetDatum(&v));
		}
		writer_close(&w);
	}
	return t;
}

int64
ref_q1_file_bytes(void *h)
{
	Q1Table    *t = h;
	int64		s = 0;

	for (int c = 0; c < NCOLS; c++)
		s += t->col[c].len;
	return s;
}

void
ref_q1_free(void *h)
{
	Q1Table    *t = h;

	if (!t)
		return;
	for (int c = 0; c < NCOLS; c++)
		free(t->col[c].bytes);
	free(t);
}

/* ---- the scan: one cursor per projected column ---- */
typedef struct ColReader
{
	ColFile    *f;
	DatumStreamBlockRead br;
	int64		pos;
	int			checksum;
	int			left;			/* rows not yet advanced over in the current block */
} ColReader;

/* next storage block of the column file: header, checksums, GetReady (datumstreamread_block, datumstream.c:1364) */
static bool
reader_next_block(ColReader *r)
{
	const int	version = AOSegfileFormatVersion_GetLatest();
	uint8	   *hdr;
	AOHeaderCheckError e;
	AoHeaderKind kind = 0;
	int32		hlen = 0,
				overall = 0,
				offset = 0,
				uncompressed = 0,
				compressed = 0;
	int			exec_kind = 0,
				rowcnt = 0;
	bool		has_first = false,
				is_compressed = false,
				adjusted = false;
	int32		adjusted_count = 0;
	int64		first_row = -1;
	RelFileNode node;

	if (r->pos >= r->f->len)
		return false;
	hdr = r->f->bytes + r->pos;
	e = AppendOnlyStorageFormat_GetHeaderInfo(hdr, r->checksum != 0, &kind, &hlen);
	if (e != AOHeaderCheckOk || kind != AoHeaderKind_SmallContent)
		ref_exec_abort("unexpected storage block header");
	e = AppendOnlyStorageFormat_GetSmallContentHeaderInfo(hdr, hlen, r->checksum != 0, 1 << 21, &overall, &offset, &uncompressed,
														  &exec_kind, &has_first, version, &first_row, &rowcnt, &is_compressed,
														  &compressed);
	if (e != AOHeaderCheckOk || is_compressed)
		ref_exec_abort("bad small content header");
	if (r->checksum)
	{
		pg_crc32	stored,
					computed;

		if (!AppendOnlyStorageFormat_VerifyHeaderChecksum(hdr, &stored, &computed) ||
			!AppendOnlyStorageFormat_VerifyBlockChecksum(hdr, overall, &stored, &computed))
			ref_exec_abort("block checksum mismatch");
	}
	memset(&node, 0, sizeof(node));
	DatumStreamBlockRead_Reset(&r->br);
	DatumStreamBlockRead_GetReady(&r->br, hdr + offset, uncompressed, first_row, rowcnt, &adjusted, &adjusted_count, &node);
	r->left = adjusted ? adjusted_count : rowcnt;
	r->pos += overall;
	return true;
}

static void
reader_open(ColReader *r, ColFile *f, int checksum)
{
	memset(r, 0, sizeof(*r));
	r->f = f;
	r->checksum = checksum;
	DatumStreamBlockRead_Init(&r->br, &f->ti, DatumStreamVersion_Original, false, NULL, NULL, NULL, NULL);
}

/* datumstreamread_advance + datumstreamread_get (datumstream.h:276-330) */
static inline bool
reader_next(ColReader *r, Datum *d, bool *isnull)
{
	if (r->left == 0 && !reader_next_block(r))
		return false;
	if (DatumStreamBlockRead_Advance(&r->br) == 0)
	{
		static char msg[160];

		snprintf(msg, sizeof(msg), "block ended early (nth %d of %d, file pos %lld, left %d, typid %d)", r->br.nth, r->br.logical_row_count, (long long) r->pos, r->left, (int) r->f->ti.typid);
		ref_exec_abort(msg);
	}
	r->left--;
	DatumStreamBlockRead_Get(&r->br, d, isnull);
	return true;
}

/* ---- grouping and aggregation ---- */
#define NSTATES 5
enum { S_QTY, S_EXT, S_DISC_PRICE, S_CHARGE, S_DISC };

typedef struct Group
{
	bool		used;
	uint32		hash;
	struct varlena *key[2];
	Datum		state[NSTATES];
	bool		state_null[NSTATES];
	int64		count;
} Group;

#define NSLOTS 64

static inline Datum
call1(PGFunction f, Oid coll, Datum a)
{
	LOCAL_FCINFO(fcinfo, 1);

	InitFunctionCallInfoData(*fcinfo, NULL, 1, coll, NULL, NULL);
	fcinfo->args[0].value = a;
	fcinfo->args[0].isnull = false;
	return (*f) (fcinfo);
}

static inline Datum
call2(PGFunction f, Oid coll, Datum a, Datum b)
{
	LOCAL_FCINFO(fcinfo, 2);

	InitFunctionCallInfoData(*fcinfo, NULL, 2, coll, NULL, NULL);
	fcinfo->args[0].value = a;
	fcinfo->args[0].isnull = false;
	fcinfo->args[1].value = b;
	fcinfo->args[1].isnull = false;
	return (*f) (fcinfo);
}

/* advance_transition_function (nodeAgg.c:725) for a non-strict transition function with an internal state */
static inline void
advance(Group *g, int s, Datum arg)
{
	LOCAL_FCINFO(fcinfo, 2);

	InitFunctionCallInfoData(*fcinfo, NULL, 2, InvalidOid, NULL, NULL);
	fcinfo->args[0].value = g->state[s];
	fcinfo->args[0].isnull = g->state_null[s];
	fcinfo->args[1].value = arg;
	fcinfo->args[1].isnull = false;
	g->state[s] = numeric_avg_accum(fcinfo);
	g->state_null[s] = fcinfo->isnull;
}

static void
final_text(PGFunction f, Group *g, int s, char *out, int cap)
{
	LOCAL_FCINFO(fcinfo, 1);
	Datum		r;

	InitFunctionCallInfoData(*fcinfo, NULL, 1, InvalidOid, NULL, NULL);
	fcinfo->args[0].value = g->state[s];
	fcinfo->args[0].isnull = g->state_null[s];
	r = (*f) (fcinfo);
	if (fcinfo->isnull)
		out[0] = 0;
	else
		snprintf(out, (size_t) cap, "%s", DatumGetCString(call1(numeric_out, InvalidOid, r)));
}

static int
group_cmp(const void *a, const void *b)
{
	const Group *x = *(Group *const *) a,
			   *y = *(Group *const *) b;

	for (int k = 0; k < 2; k++)
	{
		int			c = (int) (unsigned char) VARDATA_ANY(x->key[k])[0] - (int) (unsigned char) VARDATA_ANY(y->key[k])[0];

		if (c)
			return c;
	}
	return 0;
}

/*
 * select l_returnflag, l_linestatus, sum(l_quantity), sum(l_extendedprice), sum(l_extendedprice * (1 - l_discount)),
 *        sum(l_extendedprice * (1 - l_discount) * (1 + l_tax)), avg(l_quantity), avg(l_extendedprice), avg(l_discount), count(*)
 * from lineitem where l_shipdate <= :cutoff group by 1, 2 order by 1, 2          (rpt_tpch.source:346-371)
 * out: one line per group, '|' between columns.  Returns the number of groups, -1 on error.
 */
int
ref_q1_run(void *h, int32 cutoff, char *out, int cap, int64 *rows_passed)
{
	Q1Table    *t = h;
	ColReader	rd[NCOLS];
	Group		groups[NSLOTS];
	Group	   *order[NSLOTS];
	int			ngroups = 0;
	int64		passed = 0;
	Datum		one;
	int			pos = 0;

	if (setjmp(*ref_exec_jmp()))
	{
		ref_arena_enable(0);
		return -1;
	}
	memset(groups, 0, sizeof(groups));
	ref_arena_enable(0);
	one = DirectFunctionCall3Coll(numeric_in, InvalidOid, CStringGetDatum("1"), ObjectIdGetDatum(InvalidOid), Int32GetDatum(-1));
	for (int c = 0; c < NCOLS; c++)
		reader_open(&rd[c], &t->col[c], t->checksum);
	ref_arena_enable(1);

	for (int64 row = 0; row < t->nrows; row++)
	{
		Datum		v[NCOLS];
		bool		isnull[NCOLS];
		uint32		hashkey = 0;
		Group	   *g;
		Datum		one_minus_disc,
					disc_price,
					charge;

		ref_arena_reset();		/* ResetExprContext(econtext) per tuple */
		/* aocs_getnext: every projected column's cursor advances for every row (for a row the pushed-down qual rejects the
		 * reference skips the remaining columns' Get, aocsam.c:1359-1360; Q1's qual rejects under 2 % of the rows) */
		for (int c = 0; c < NCOLS; c++)
			if (!reader_next(&rd[c], &v[c], &isnull[c]))
				ref_exec_abort("column file ended early");
		/* the pushed-down qual: l_shipdate <= cutoff (date_le on DateADT) */
		if (DatumGetInt32(v[C_SHIPDATE]) > cutoff)
			continue;
		passed++;

		/* TupleHashTableHash_internal (execGrouping.c:443-495): hash_iv = 0, rotate-xor of the key hashes, murmurhash32 */
		for (int k = 0; k < 2; k++)
		{
			hashkey = (hashkey << 1) | ((hashkey & 0x80000000) ? 1 : 0);
			hashkey ^= DatumGetUInt32(call1(hashbpchar, DEFAULT_COLLATION_OID, v[C_RF + k]));
		}
		hashkey = murmurhash32(hashkey);
		for (uint32 slot = hashkey & (NSLOTS - 1);; slot = (slot + 1) & (NSLOTS - 1))
		{
			g = &groups[slot];
			if (!g->used)
			{
				/* new group: the key tuple is copied into the table's context (execGrouping.c:533) */
				if (++ngroups > NSLOTS / 2)
					ref_exec_abort("more groups than this driver's table holds");
				g->used = true;
				g->hash = hashkey;
				for (int k = 0; k < 2; k++)
				{
					struct varlena *src = (struct varlena *) DatumGetPointer(v[C_RF + k]);
					Size		len = VARSIZE_ANY(src);

					g->key[k] = malloc(len);
					memcpy(g->key[k], src, len);
				}
				for (int s = 0; s < NSTATES; s++)
					g->state_null[s] = true;
				break;
			}
			if (g->hash == hashkey &&
				DatumGetBool(call2(bpchareq, DEFAULT_COLLATION_OID, PointerGetDatum(g->key[0]), v[C_RF])) &&
				DatumGetBool(call2(bpchareq, DEFAULT_COLLATION_OID, PointerGetDatum(g->key[1]), v[C_LS])))
				break;
		}

		/* the aggregates' argument expressions; the planner does not share the common sub-expression, so neither does this */
		one_minus_disc = call2(numeric_sub, InvalidOid, one, v[C_DISC]);
		disc_price = call2(numeric_mul, InvalidOid, v[C_EXT], one_minus_disc);
		one_minus_disc = call2(numeric_sub, InvalidOid, one, v[C_DISC]);
		charge = call2(numeric_mul, InvalidOid,
					   call2(numeric_mul, InvalidOid, v[C_EXT], one_minus_disc),
					   call2(numeric_add, InvalidOid, one, v[C_TAX]));
		advance(g, S_QTY, v[C_QTY]);	/* sum(l_quantity), avg(l_quantity) */
		advance(g, S_EXT, v[C_EXT]);	/* sum(l_extendedprice), avg(l_extendedprice) */
		advance(g, S_DISC_PRICE, disc_price);
		advance(g, S_CHARGE, charge);
		advance(g, S_DISC, v[C_DISC]);	/* avg(l_discount) */
		g->count++;				/* int8inc */
	}
	for (int c = 0; c < NCOLS; c++)
		DatumStreamBlockRead_Finish(&rd[c].br);

	/* finalize_aggregates + the Sort above the Agg */
	ngroups = 0;
	for (int i = 0; i < NSLOTS; i++)
		if (groups[i].used)
			order[ngroups++] = &groups[i];
	qsort(order, (size_t) ngroups, sizeof(order[0]), group_cmp);
	ref_arena_enable(0);
	out[0] = 0;
	for (int i = 0; i < ngroups; i++)
	{
		Group	   *g = order[i];
		char		f[8][96];

		final_text(numeric_sum, g, S_QTY, f[0], 96);
		final_text(numeric_sum, g, S_EXT, f[1], 96);
		final_text(numeric_sum, g, S_DISC_PRICE, f[2], 96);
		final_text(numeric_sum, g, S_CHARGE, f[3], 96);
		final_text(numeric_avg, g, S_QTY, f[4], 96);
		final_text(numeric_avg, g, S_EXT, f[5], 96);
		final_text(numeric_avg, g, S_DISC, f[6], 96);
		pos += snprintf(out + pos, pos < cap ? (size_t) (cap - pos) : 0, "%c|%c|%s|%s|%s|%s|%s|%s|%s|%lld\n",
						VARDATA_ANY(g->key[0])[0], VARDATA_ANY(g->key[1])[0], f[0], f[1], f[2], f[3], f[4], f[5], f[6],
						(long long) g->count);
		free(g->key[0]);
		free(g->key[1]);
	}
	if (rows_passed)
		*rows_passed = passed;
	return pos < cap ? ngroups : -1;
}
