/*
 * oracle/ref_exec.c - drives the REFERENCE's own per-row arithmetic of the executor hot path, so that the oracle's
 * restatements (oracle/pg_hash.h, oracle.c) and the product's host finalisation (csrc/exec/cb_numeric.c) are pinned
 * against reference code, not against a second restatement.
 *
 * TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile (only where /root/reference exists) into
 * oracle/_ref/libexec_ref.so together with these reference sources compiled where they lie:
 *     src/common/hashfn.c                     hash_bytes / hash_bytes_uint32
 *     src/backend/access/hash/hashfunc.c      hashint4, hashint8, hashfloat8, hashtext
 *     src/backend/utils/adt/varchar.c         hashbpchar (bcTruelen)
 *     src/backend/cdb/cdbhash.c               makeCdbHash, cdbhashinit, cdbhash, cdbhashreduce, jump_consistent_hash
 *     src/backend/utils/adt/numeric.c         numeric_in/out, numeric_mul/add/sub, numeric_avg_accum, int8_avg_accum,
 *                                             int4_sum, numeric_sum / numeric_avg / numeric_poly_sum / numeric_poly_avg
 *     (and, for oracle/ref_q1.c, the AOCS block reader / writer: datumstreamblock.c, cdbappendonlystorageformat.c, pg_crc32c_sb8.c)
 * The generated headers are stand-ins (oracle/ref_shim/) or are derived at build time from the reference's own
 * errcodes.txt and function definitions (oracle/gen_ref_headers.py -> oracle/_ref/gen/).  No reference source is copied:
 * this file stubs the backend services those sources call (palloc, ereport, the fmgr call helpers, interrupt flags)
 * and calls the functions the way the executor does (nodeAgg.c advance_transition_function / finalize_aggregate,
 * nodeMotion.c evalHashKey).
 */
#include "postgres.h"

#include <setjmp.h>
#include <stdarg.h>
#include <signal.h>

#include "catalog/pg_collation.h"
#include "cdb/cdbhash.h"
#include "fmgr.h"
#include "utils/builtins.h"
#include "utils/fmgrprotos.h"
#include "utils/memutils.h"
#include "utils/numeric.h"

#undef vsnprintf
#undef snprintf
#undef vsprintf
#undef sprintf
#undef printf
#undef fprintf
#undef vfprintf

/* ---- backend services ---- */
MemoryContext CurrentMemoryContext = NULL;
MemoryContext TopMemoryContext = NULL;
volatile sig_atomic_t InterruptPending = 0;
int			backoffTickCounter = 0;
int			gp_resqueue_priority_local_interval = 0x7fffffff;
bool		gp_mp_inited = false;
bool		trace_sort = false;
uint32		magic_hash_stash = 0;
/* ReportOOMConsumption (miscadmin.h) stays quiet while these are equal */
volatile OOMTimeType *segmentOOMTime = NULL;
volatile OOMTimeType oomTrackerStartTime = 0;
volatile OOMTimeType alreadyReportedOOMTime = 0;

static jmp_buf ref_jmp;
static char ref_errbuf[512];
static int	ref_elevel;

void		ref_exec_abort(const char *what);
jmp_buf    *ref_exec_jmp(void);

/* the unwind target of ereport(ERROR) / ref_exec_abort, for the drivers in ref_q1.c */
jmp_buf *
ref_exec_jmp(void)
{
	return &ref_jmp;
}

void
ref_exec_abort(const char *what)
{
	snprintf(ref_errbuf, sizeof(ref_errbuf), "oracle/_ref: %s is a stub", what);
	longjmp(ref_jmp, 1);
}

/*
 * palloc: malloc by default.  ref_q1.c switches on a bump arena that stands in for the executor's per-tuple memory context
 * (ResetExprContext once per row, execScan.c:195): allocations made while CurrentMemoryContext is REF_AGG_CONTEXT - the
 * aggregate transition states, which nodeAgg.c keeps in aggcontext - stay on malloc and survive the reset.
 */
#define REF_AGG_CONTEXT ((MemoryContext) (uintptr_t) 16)
#define ARENA_CAP ((size_t) 1 << 22)
static char *arena;
static size_t arena_off;
static int	arena_on;

void		ref_arena_enable(int on);
void		ref_arena_reset(void);

void
ref_arena_enable(int on)
{
	if (on && !arena)
		arena = malloc(ARENA_CAP);
	arena_on = on;
	arena_off = 0;
}

void		ref_arena_reset(void) { arena_off = 0; }

static inline bool
in_arena(const void *p)
{
	return arena && (const char *) p >= arena && (const char *) p < arena + ARENA_CAP;
}

void *
palloc(Size size)
{
	size_t		need = ((size ? size : 1) + 15) & ~(size_t) 15;

	if (arena_on && CurrentMemoryContext != REF_AGG_CONTEXT && arena_off + need + 16 <= ARENA_CAP)
	{
		char	   *p = arena + arena_off;

		*(size_t *) p = size;	/* repalloc needs the old size */
		arena_off += need + 16;
		return p + 16;
	}
	return malloc(size ? size : 1);
}

void *
palloc0(Size size)
{
	void	   *p = palloc(size);

	memset(p, 0, size ? size : 1);
	return p;
}

void *
repalloc(void *p, Size size)
{
	if (in_arena(p))
	{
		size_t		old = *(size_t *) ((char *) p - 16);
		void	   *q = palloc(size);

		memcpy(q, p, old < size ? old : size);
		return q;
	}
	return realloc(p, size ? size : 1);
}

void
pfree(void *p)
{
	if (!in_arena(p))
		free(p);
}

char *
pstrdup(const char *s)
{
	char	   *r = palloc(strlen(s) + 1);

	strcpy(r, s);
	return r;
}

bool
errstart(int elevel, const char *domain)
{
	(void) domain;
	ref_elevel = elevel;
	return elevel >= ERROR;
}

bool		errstart_cold(int elevel, const char *domain) { return errstart(elevel, domain); }

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
void		errhint(const char *fmt,...) { (void) fmt; }
void		errcode(int sqlerrcode) { (void) sqlerrcode; }

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

char *
psprintf(const char *fmt,...)
{
	char	   *buf = palloc(1024);
	va_list		ap;

	va_start(ap, fmt);
	vsnprintf(buf, 1024, fmt, ap);
	va_end(ap);
	return buf;
}

void		errdetail_internal(const char *fmt,...) { (void) fmt; }
int			errprintstack(bool printstack) { (void) printstack; return 0; }

/* src/port/qsort.c's name for it (port.h maps qsort to pg_qsort: undo that here or this calls itself) */
#undef qsort
void
pg_qsort(void *base, size_t nel, size_t elsize, int (*cmp) (const void *, const void *))
{
	qsort(base, nel, elsize, cmp);
}

/* fmgr.c:793-860 call helpers: one fcinfo on the stack, NULL result is an error */
Datum
DirectFunctionCall1Coll(PGFunction func, Oid collation, Datum arg1)
{
	LOCAL_FCINFO(fcinfo, 1);
	Datum		result;

	InitFunctionCallInfoData(*fcinfo, NULL, 1, collation, NULL, NULL);
	fcinfo->args[0].value = arg1;
	fcinfo->args[0].isnull = false;
	result = (*func) (fcinfo);
	if (fcinfo->isnull)
		ref_exec_abort("DirectFunctionCall1Coll: NULL result");
	return result;
}

Datum
DirectFunctionCall2Coll(PGFunction func, Oid collation, Datum arg1, Datum arg2)
{
	LOCAL_FCINFO(fcinfo, 2);
	Datum		result;

	InitFunctionCallInfoData(*fcinfo, NULL, 2, collation, NULL, NULL);
	fcinfo->args[0].value = arg1;
	fcinfo->args[0].isnull = false;
	fcinfo->args[1].value = arg2;
	fcinfo->args[1].isnull = false;
	result = (*func) (fcinfo);
	if (fcinfo->isnull)
		ref_exec_abort("DirectFunctionCall2Coll: NULL result");
	return result;
}

Datum
DirectFunctionCall3Coll(PGFunction func, Oid collation, Datum arg1, Datum arg2, Datum arg3)
{
	LOCAL_FCINFO(fcinfo, 3);
	Datum		result;

	InitFunctionCallInfoData(*fcinfo, NULL, 3, collation, NULL, NULL);
	fcinfo->args[0].value = arg1;
	fcinfo->args[0].isnull = false;
	fcinfo->args[1].value = arg2;
	fcinfo->args[1].isnull = false;
	fcinfo->args[2].value = arg3;
	fcinfo->args[2].isnull = false;
	result = (*func) (fcinfo);
	if (fcinfo->isnull)
		ref_exec_abort("DirectFunctionCall3Coll: NULL result");
	return result;
}

/* the datums these drivers pass are plain 4-byte-header or 1-byte-header varlenas: never toasted, never compressed */
struct varlena *
pg_detoast_datum(struct varlena *datum)
{
	if (VARATT_IS_SHORT(datum))
	{
		Size		data_size = VARSIZE_SHORT(datum) - VARHDRSZ_SHORT;
		struct varlena *res = palloc(data_size + VARHDRSZ);

		SET_VARSIZE(res, data_size + VARHDRSZ);
		memcpy(VARDATA(res), VARDATA_SHORT(datum), data_size);
		return res;
	}
	return datum;
}

struct varlena *pg_detoast_datum_packed(struct varlena *datum) { return datum; }

/* nodeAgg.c:4893 AggCheckCallContext: these calls are always "inside an aggregate" */
int
AggCheckCallContext(FunctionCallInfo fcinfo, MemoryContext *aggcontext)
{
	(void) fcinfo;
	if (aggcontext)
		*aggcontext = REF_AGG_CONTEXT;
	return AGG_CONTEXT_AGGREGATE;
}

bool		lc_collate_is_c(Oid collation) { return collation == C_COLLATION_OID || collation == POSIX_COLLATION_OID; }

/* fmgr_info for the hash support functions makeCdbHash loads (OIDs: catalog/pg_proc.dat) */
#define F_HASHINT4_OID 450
#define F_HASHINT8_OID 949
#define F_HASHFLOAT8_OID 452
#define F_HASHTEXT_OID 400
#define F_HASHBPCHAR_OID 1080

void
fmgr_info(Oid functionId, FmgrInfo *finfo)
{
	memset(finfo, 0, sizeof(*finfo));
	finfo->fn_oid = functionId;
	finfo->fn_nargs = 1;
	finfo->fn_strict = true;
	switch (functionId)
	{
		case F_HASHINT4_OID: finfo->fn_addr = hashint4; break;
		case F_HASHINT8_OID: finfo->fn_addr = hashint8; break;
		case F_HASHFLOAT8_OID: finfo->fn_addr = hashfloat8; break;
		case F_HASHTEXT_OID: finfo->fn_addr = hashtext; break;
		case F_HASHBPCHAR_OID: finfo->fn_addr = hashbpchar; break;
		default: ref_exec_abort("fmgr_info: function not in the table");
	}
}

bool		isLegacyCdbHashFunction(Oid funcid) { (void) funcid; return false; }

void		RedZoneHandler_DetectRunawaySession(void) { }

/* numeric_in compares "NaN" / "Infinity" spellings with it (src/port/pgstrcasecmp.c:69); ASCII is all that reaches it */
int
pg_strncasecmp(const char *s1, const char *s2, size_t n)
{
	while (n-- > 0)
	{
		unsigned char a = (unsigned char) *s1++, b = (unsigned char) *s2++;

		if (a >= 'A' && a <= 'Z') a += 'a' - 'A';
		if (b >= 'A' && b <= 'Z') b += 'a' - 'A';
		if (a != b)
			return (int) a - (int) b;
		if (a == 0)
			break;
	}
	return 0;
}

/* ---------------------------------------------------------------- drivers ---------------------------------------------------------------- */

const char *ref_exec_last_error(void);
uint32		ref_hash_datum(int kind, int64 bits, const char *bytes, int len);
int			ref_cdbhash_segment(int natts, const int *kinds, const int64 *bits, const char *const *strs, const uint8 *isnull, int numsegs);
int			ref_jump_consistent_hash(uint32 key, int nbuckets);
int			ref_numeric_binop(int op, const char *a, const char *b, char *out, int cap);
int			ref_numeric_agg(const char *const *vals, int n, char *sum_out, char *avg_out, int cap);
int			ref_int8_agg(const int64 *vals, int n, char *sum_out, char *avg_out, int cap);
int			ref_int4_sum(const int32 *vals, int n, int64 *out);

const char *
ref_exec_last_error(void)
{
	return ref_errbuf;
}

/* a varlena with a 4-byte header around len bytes */
static struct varlena *
make_varlena(const char *bytes, int len)
{
	struct varlena *v = malloc((size_t) len + VARHDRSZ);

	SET_VARSIZE(v, len + VARHDRSZ);
	memcpy(VARDATA(v), bytes, (size_t) len);
	return v;
}

enum { REF_INT4 = 0, REF_INT8 = 1, REF_FLOAT8 = 2, REF_BPCHAR = 3, REF_TEXT = 4 };

static Oid
hash_proc_of(int kind)
{
	switch (kind)
	{
		case REF_INT4: return F_HASHINT4_OID;
		case REF_INT8: return F_HASHINT8_OID;
		case REF_FLOAT8: return F_HASHFLOAT8_OID;
		case REF_BPCHAR: return F_HASHBPCHAR_OID;
		case REF_TEXT: return F_HASHTEXT_OID;
	}
	ref_exec_abort("hash kind");
	return 0;
}

static Datum
datum_of(int kind, int64 bits, const char *bytes, int len)
{
	switch (kind)
	{
		case REF_INT4: return Int32GetDatum((int32) bits);
		case REF_INT8: return Int64GetDatum(bits);
		case REF_FLOAT8: return (Datum) bits;	/* Float8GetDatum of the same bit pattern (by value, 8 bytes) */
		default: return PointerGetDatum(make_varlena(bytes, len));
	}
}

/* the type's hash support function as the executor calls it for joins / groupings (nodeHash.c:2146, execGrouping.c:483) */
uint32
ref_hash_datum(int kind, int64 bits, const char *bytes, int len)
{
	FmgrInfo	fi;

	if (setjmp(ref_jmp))
		return 0xdeadbeef;
	fmgr_info(hash_proc_of(kind), &fi);
	return DatumGetUInt32(DirectFunctionCall1Coll(fi.fn_addr, DEFAULT_COLLATION_OID, datum_of(kind, bits, bytes, len)));
}

/* evalHashKey (nodeMotion.c:1088): cdbhashinit, cdbhash per key, cdbhashreduce; -1 on error */
int
ref_cdbhash_segment(int natts, const int *kinds, const int64 *bits, const char *const *strs, const uint8 *isnull, int numsegs)
{
	CdbHash    *h;
	Oid			procs[16];
	int			seg;

	if (setjmp(ref_jmp))
		return -1;
	if (natts > 16)
		ref_exec_abort("too many keys");
	for (int i = 0; i < natts; i++)
		procs[i] = hash_proc_of(kinds[i]);
	h = makeCdbHash(numsegs, natts, procs);
	cdbhashinit(h);
	for (int i = 0; i < natts; i++)
	{
		bool		nul = isnull && isnull[i];

		cdbhash(h, i + 1, nul ? (Datum) 0 : datum_of(kinds[i], bits[i], strs ? strs[i] : NULL, strs && strs[i] ? (int) strlen(strs[i]) : 0), nul);
	}
	seg = (int) cdbhashreduce(h);
	return seg;
}

/* jump_consistent_hash (cdbhash.c:530, static) through its only caller: cdbhashreduce of a preset 32-bit hash value */
int
ref_jump_consistent_hash(uint32 key, int nbuckets)
{
	CdbHash    *h;
	Oid			proc = F_HASHINT4_OID;

	if (setjmp(ref_jmp))
		return -1;
	h = makeCdbHash(nbuckets, 1, &proc);
	h->hash = key;
	return (int) cdbhashreduce(h);
}

static Datum
num_in(const char *s)
{
	return DirectFunctionCall3Coll(numeric_in, InvalidOid, CStringGetDatum(s), ObjectIdGetDatum(InvalidOid), Int32GetDatum(-1));
}

static void
num_out(Datum d, char *out, int cap)
{
	char	   *s = DatumGetCString(DirectFunctionCall1Coll(numeric_out, InvalidOid, d));

	snprintf(out, (size_t) cap, "%s", s);
}

/* op 0 = numeric_add, 1 = numeric_sub, 2 = numeric_mul: what ExecInterpExpr calls for l_extendedprice * (1 - l_discount);
 * 3 = numeric_div, which is all numeric_avg does with (sumX, N) (numeric.c:6056-6088) */
int
ref_numeric_binop(int op, const char *a, const char *b, char *out, int cap)
{
	PGFunction	f = op == 0 ? numeric_add : op == 1 ? numeric_sub : op == 2 ? numeric_mul : numeric_div;

	if (setjmp(ref_jmp))
		return -1;
	num_out(DirectFunctionCall2Coll(f, InvalidOid, num_in(a), num_in(b)), out, cap);
	return 0;
}

/* one transition call as advance_transition_function makes it (nodeAgg.c:725): arg0 = state (NULL before the first row) */
static Datum
trans_call(PGFunction f, Datum state, bool state_null, Datum arg, bool *res_null)
{
	LOCAL_FCINFO(fcinfo, 2);
	Datum		r;

	InitFunctionCallInfoData(*fcinfo, NULL, 2, InvalidOid, NULL, NULL);
	fcinfo->args[0].value = state;
	fcinfo->args[0].isnull = state_null;
	fcinfo->args[1].value = arg;
	fcinfo->args[1].isnull = false;
	r = (*f) (fcinfo);
	*res_null = fcinfo->isnull;
	return r;
}

/* a final function over the state (finalize_aggregate, nodeAgg.c:1115); "" in out for a NULL result */
static void
final_call(PGFunction f, Datum state, bool state_null, char *out, int cap)
{
	LOCAL_FCINFO(fcinfo, 1);
	Datum		r;

	InitFunctionCallInfoData(*fcinfo, NULL, 1, InvalidOid, NULL, NULL);
	fcinfo->args[0].value = state;
	fcinfo->args[0].isnull = state_null;
	r = (*f) (fcinfo);
	if (fcinfo->isnull)
		out[0] = 0;
	else
		num_out(r, out, cap);
}

/* sum(numeric) / avg(numeric): numeric_avg_accum over every value, then numeric_sum and numeric_avg (pg_aggregate.dat:33,81) */
int
ref_numeric_agg(const char *const *vals, int n, char *sum_out, char *avg_out, int cap)
{
	Datum		state = (Datum) 0;
	bool		isnull = true;

	if (setjmp(ref_jmp))
		return -1;
	for (int i = 0; i < n; i++)
		state = trans_call(numeric_avg_accum, state, isnull, num_in(vals[i]), &isnull);
	final_call(numeric_sum, state, isnull, sum_out, cap);
	final_call(numeric_avg, state, isnull, avg_out, cap);
	return 0;
}

/* sum(int8) / avg(int8): int8_avg_accum, then numeric_poly_sum and numeric_poly_avg (pg_aggregate.dat:16,55) */
int
ref_int8_agg(const int64 *vals, int n, char *sum_out, char *avg_out, int cap)
{
	Datum		state = (Datum) 0;
	bool		isnull = true;

	if (setjmp(ref_jmp))
		return -1;
	for (int i = 0; i < n; i++)
		state = trans_call(int8_avg_accum, state, isnull, Int64GetDatum(vals[i]), &isnull);
	final_call(numeric_poly_sum, state, isnull, sum_out, cap);
	final_call(numeric_poly_avg, state, isnull, avg_out, cap);
	return 0;
}

/* sum(int4) -> int8: int4_sum is its own transition function, no final function (pg_aggregate.dat:62) */
int
ref_int4_sum(const int32 *vals, int n, int64 *out)
{
	Datum		state = (Datum) 0;
	bool		isnull = true;

	if (setjmp(ref_jmp))
		return -1;
	for (int i = 0; i < n; i++)
		state = trans_call(int4_sum, state, isnull, Int32GetDatum(vals[i]), &isnull);
	if (isnull)
		return 1;
	*out = DatumGetInt64(state);
	return 0;
}

/* ------------------------------------------------------------------------------------------
 * partial aggregate states as the reference ships them between a Partial and a Finalize stage (serialfn / deserialfn of
 * pg_aggregate.dat: numeric_avg_serialize / numeric_avg_deserialize for sum / avg over numeric, int8_avg_serialize /
 * int8_avg_deserialize for sum / avg over int8) - the checker of cb_numeric_avg_serialize / cb_int8_avg_serialize /
 * cb_numeric_avg_deserialize (csrc/exec/cb_numeric.c).
 * ------------------------------------------------------------------------------------------ */
static int
bytea_out(Datum d, unsigned char *out, int cap)
{
	struct varlena *v = pg_detoast_datum((struct varlena *) DatumGetPointer(d));
	int			len = (int) VARSIZE_ANY_EXHDR(v);

	if (len > cap)
		return -2;
	memcpy(out, VARDATA_ANY(v), len);
	return len;
}

static Datum
unary_call(PGFunction f, Datum arg)
{
	LOCAL_FCINFO(fcinfo, 2);

	InitFunctionCallInfoData(*fcinfo, NULL, 2, InvalidOid, NULL, NULL);
	fcinfo->args[0].value = arg;
	fcinfo->args[0].isnull = false;
	fcinfo->args[1].value = (Datum) 0;	/* the deserialisation functions' dummy second argument */
	fcinfo->args[1].isnull = false;
	return (*f) (fcinfo);
}

/* Partial stage on the CPU: accumulate, then the serialisation function's bytea (its payload, without the varlena header) */
int
ref_numeric_avg_serialize(const char *const *vals, int n, unsigned char *out, int cap)
{
	Datum		state = (Datum) 0;
	bool		isnull = true;

	if (setjmp(ref_jmp))
		return -1;
	for (int i = 0; i < n; i++)
		state = trans_call(numeric_avg_accum, state, isnull, num_in(vals[i]), &isnull);
	if (isnull)
		return -3;
	return bytea_out(unary_call(numeric_avg_serialize, state), out, cap);
}

int
ref_int8_avg_serialize(const int64 *vals, int n, unsigned char *out, int cap)
{
	Datum		state = (Datum) 0;
	bool		isnull = true;

	if (setjmp(ref_jmp))
		return -1;
	for (int i = 0; i < n; i++)
		state = trans_call(int8_avg_accum, state, isnull, Int64GetDatum(vals[i]), &isnull);
	if (isnull)
		return -3;
	return bytea_out(unary_call(int8_avg_serialize, state), out, cap);
}

static Datum
make_bytea(const unsigned char *bytes, int len)
{
	struct varlena *v = (struct varlena *) palloc(VARHDRSZ + len);

	SET_VARSIZE(v, VARHDRSZ + len);
	memcpy(VARDATA(v), bytes, len);
	return PointerGetDatum(v);
}

/* Finalize stage on the CPU over a state somebody else serialised: deserialise, then the final functions (sum and avg as text) */
int
ref_numeric_avg_finalize(const unsigned char *bytes, int len, char *sum_out, char *avg_out, int cap)
{
	Datum		state;

	if (setjmp(ref_jmp))
		return -1;
	state = unary_call(numeric_avg_deserialize, make_bytea(bytes, len));
	final_call(numeric_sum, state, false, sum_out, cap);
	final_call(numeric_avg, state, false, avg_out, cap);
	return 0;
}

int
ref_int8_avg_finalize(const unsigned char *bytes, int len, char *sum_out, char *avg_out, int cap)
{
	Datum		state;

	if (setjmp(ref_jmp))
		return -1;
	state = unary_call(int8_avg_deserialize, make_bytea(bytes, len));
	final_call(numeric_poly_sum, state, false, sum_out, cap);
	final_call(numeric_poly_avg, state, false, avg_out, cap);
	return 0;
}
