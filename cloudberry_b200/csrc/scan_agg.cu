/*
 * scan_agg.cu - K1+K4 fused: columnar scan -> qual -> low-cardinality hash aggregate in one pass
 * (the TPC-H Q1 class: few groups keyed by one or two char(1)/dictionary columns, SUM/AVG/COUNT
 * over int64-scaled numerics and products of the form b*(k-c) and b*(k-c)*(k2+d)).
 *
 * What it fuses, in the reference: aocs_getnext over the projected columns (backend/access/aocs/
 * aocsam.c:1418,1138), the pushed-down scan qual (aocsam.c:1269, execScan.c:162), the numeric
 * arithmetic of the aggregate arguments (numeric_mul / numeric_sub, backend/utils/adt/numeric.c:
 * 2645,2567 - here exact scaled-integer products), LookupTupleHashEntry (backend/executor/
 * execGrouping.c:317) and the transition functions (numeric_avg_accum, int8inc: nodeAgg.c:856).
 *
 * Shape: persistent grid (a multiple of the SM count), each thread streams rows with 16-byte
 * vectorised loads (two rows per load), keeps per-group accumulators in registers (G <= 4 or 8
 * groups, slot ids from a per-CTA shared-memory key table), reduces per CTA with warp shuffles and
 * commits each group once per CTA into the global aggregate table with 128-bit exact adds.
 * HBM-bound: algorithmic bytes/row = sum of the projected column widths (Q1: 38 B/row).
 *
 * Exactness: per-row products are 64-bit; the kernel ORs the magnitudes of the factors it
 * multiplies and of the values it accumulates, and reports CBGPU_ERR_OVERFLOW (never a wrapped
 * sum) if a product or a per-thread partial sum could have left 63 bits.  Group totals are 128-bit.
 */
#include "pipeline.cuh"
#include "xmatch.h"

#define SA_THREADS 256
#define SA_NSUM 6				/* sA sB sC sD sREV sCHG */

struct SmallAggParams
{
	int64_t		nrows;
	const long long *colA;		/* optional plain column                                              */
	const long long *colB;		/* b of b*(k-c)                                                       */
	const long long *colC;
	const long long *colD;		/* optional: d of (k2+d)                                              */
	long long	k,
				k2;
	const int32_t *fcol;		/* optional filter column (int4/date)                                 */
	int32_t		fcode;			/* CbpOpCode comparison                                               */
	int32_t		fconst;
	const uint8_t *key0;		/* optional group key columns (char(1) / dictionary code)             */
	const uint8_t *key1;
	const uint8_t *visimap;
	/* commit: agg accumulator a takes sum index accsum[a] (0..5) or -1 for a pure count */
	AggDev		agg;
	int32_t		nkeys;
	int32_t		keytype[2];
	const uint32_t *keydict[2];
	int32_t		naccs;
	int32_t		accsum[CBP_MAX_AGGS];
	int32_t		want_chg;		/* the b*(k-c)*(k2+d) sum is requested                                */
	int		   *status;
	int		   *retry;			/* set when a CTA saw more than G distinct groups                     */
	/* per-CTA partial results, committed by k_small_commit once no CTA asked for a retry */
	unsigned long long *scratch;	/* [grid][G][SA_NSUM + 1][2]                                      */
	unsigned   *skeys;			/* [grid][G]                                                          */
};

__device__ __forceinline__ bool
sa_cmp(int code, int32_t x, int32_t y)
{
	switch (code)
	{
		case CBP_EQ: return x == y;
		case CBP_NE: return x != y;
		case CBP_LT: return x < y;
		case CBP_LE: return x <= y;
		case CBP_GT: return x > y;
		default: return x >= y;
	}
}

__device__ __forceinline__ void
ld_nc_v2(const long long *p, long long &a, long long &b)
{
	asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
}

__device__ __forceinline__ long long
sa_abs(long long v)
{
	return v < 0 ? -v : v;
}

template <int G>
struct SaAcc
{
	unsigned	cnt[G];
	long long	s[G][SA_NSUM];
};

/* accumulate one row into the register accumulators of its group slot */
template <int G>
__device__ __forceinline__ void
sa_accumulate(SaAcc<G> &acc, int slot, bool pass, long long a, long long b, long long c, long long d, long long rev, long long chg)
{
#pragma unroll
	for (int g = 0; g < G; g++)
	{
		bool		hit = pass && slot == g;

		acc.cnt[g] += hit ? 1u : 0u;
		acc.s[g][0] += hit ? a : 0;
		acc.s[g][1] += hit ? b : 0;
		acc.s[g][2] += hit ? c : 0;
		acc.s[g][3] += hit ? d : 0;
		acc.s[g][4] += hit ? rev : 0;
		acc.s[g][5] += hit ? chg : 0;
	}
}

/* find (or claim) the slot of a group key in the CTA's shared key table */
template <int G>
__device__ __forceinline__ int
sa_slot(unsigned *gkeys, unsigned key, int *retry)
{
#pragma unroll
	for (int g = 0; g < G; g++)
	{
		unsigned	cur = ((volatile unsigned *) gkeys)[g];

		if (cur == key)
			return g;
		if (cur == 0)
		{
			unsigned	old = atomicCAS(gkeys + g, 0u, key);

			if (old == 0 || old == key)
				return g;
		}
	}
	atomicExch(retry, 1);
	return -1;
}

template <int G>
__global__ void __launch_bounds__(SA_THREADS, (G <= 4 ? 2 : 1))
k_scan_agg_small(const __grid_constant__ SmallAggParams P)
{
	__shared__ unsigned gkeys[G];
	__shared__ unsigned long long red[G][SA_NSUM + 1][2];	/* 128-bit CTA totals                     */
	SaAcc<G>	acc;
	unsigned long long magB = 0,
				magKC = 0,
				magKD = 0,
				magAcc = 0;
	long long	rows_seen = 0;

	if (threadIdx.x < G)
		gkeys[threadIdx.x] = 0;
	for (int i = threadIdx.x; i < G * (SA_NSUM + 1) * 2; i += blockDim.x)
		(&red[0][0][0])[i] = 0;
#pragma unroll
	for (int g = 0; g < G; g++)
	{
		acc.cnt[g] = 0;
#pragma unroll
		for (int s = 0; s < SA_NSUM; s++)
			acc.s[g][s] = 0;
	}
	__syncthreads();

	/* tiles of 2 * SA_THREADS rows: thread t owns rows base + 2t, base + 2t + 1 */
	const int64_t tile = 2 * SA_THREADS;
	const int64_t ntiles = (P.nrows + tile - 1) / tile;

	for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x)
	{
		const int64_t r0 = t * tile + 2 * threadIdx.x;
		long long	a[2] = {0, 0},
					b[2],
					c[2],
					d[2] = {0, 0};
		int32_t		f[2] = {0, 0};
		unsigned	k0[2] = {0, 0},
					k1[2] = {0, 0};
		bool		ok[2];

		ok[0] = r0 < P.nrows;
		ok[1] = r0 + 1 < P.nrows;
		if (ok[1])
		{
			/* full pair: 16-byte loads, streaming (no L1 allocation: each byte is used once) */
			ld_nc_v2(P.colB + r0, b[0], b[1]);
			ld_nc_v2(P.colC + r0, c[0], c[1]);
			if (P.colA)
				ld_nc_v2(P.colA + r0, a[0], a[1]);
			if (P.colD)
				ld_nc_v2(P.colD + r0, d[0], d[1]);
			if (P.fcol)
			{
				int2		fv = __ldg((const int2 *) (P.fcol + r0));

				f[0] = fv.x;
				f[1] = fv.y;
			}
			if (P.key0)
			{
				unsigned short kv = __ldg((const unsigned short *) (P.key0 + r0));

				k0[0] = kv & 0xff;
				k0[1] = kv >> 8;
			}
			if (P.key1)
			{
				unsigned short kv = __ldg((const unsigned short *) (P.key1 + r0));

				k1[0] = kv & 0xff;
				k1[1] = kv >> 8;
			}
		}
		else if (ok[0])
		{
			b[0] = __ldg(P.colB + r0);
			c[0] = __ldg(P.colC + r0);
			b[1] = c[1] = 0;
			if (P.colA)
				a[0] = __ldg(P.colA + r0);
			if (P.colD)
				d[0] = __ldg(P.colD + r0);
			if (P.fcol)
				f[0] = __ldg(P.fcol + r0);
			if (P.key0)
				k0[0] = __ldg(P.key0 + r0);
			if (P.key1)
				k1[0] = __ldg(P.key1 + r0);
		}
		else
		{
			b[0] = b[1] = c[0] = c[1] = 0;
		}
		if (P.visimap && ok[0])
		{
			/* AppendOnlyVisimap_IsVisible: both rows of the pair share a byte (r0 is even) */
			unsigned	vb = __ldg(P.visimap + (r0 >> 3)) >> (r0 & 7);

			ok[0] = ok[0] && (vb & 1);
			ok[1] = ok[1] && ((vb >> 1) & 1);
		}
#pragma unroll
		for (int j = 0; j < 2; j++)
		{
			bool		pass = ok[j] && (!P.fcol || sa_cmp(P.fcode, f[j], P.fconst));
			long long	kc = P.k - c[j];
			long long	kd = P.k2 + d[j];
			long long	rev = b[j] * kc;
			long long	chg = P.want_chg ? rev * kd : 0;
			/* slot lookup: the key carries a valid bit so 0 means "empty slot" */
			unsigned	key = 0x10000u | k0[j] | (k1[j] << 8);
			int			slot = 0;

			if (pass)
			{
				slot = sa_slot<G>(gkeys, key, P.retry);
				magB |= (unsigned long long) sa_abs(b[j]);
				magKC |= (unsigned long long) sa_abs(kc);
				if (P.want_chg)
					magKD |= (unsigned long long) sa_abs(kd);
				magAcc |= (unsigned long long) (sa_abs(a[j]) | sa_abs(b[j]) | sa_abs(c[j]) | sa_abs(d[j]) | sa_abs(rev) | sa_abs(chg));
				rows_seen++;
			}
			sa_accumulate<G>(acc, slot, pass && slot >= 0, a[j], b[j], c[j], d[j], rev, chg);
		}
	}

	/* overflow audit: bits(b) + bits(k-c) + bits(k2+d) must stay below 63, and so must
	 * bits(largest accumulated value) + bits(rows this thread accumulated) */
	{
		int			bb = 64 - __clzll(magB),
					bkc = 64 - __clzll(magKC),
					bkd = 64 - __clzll(magKD);
		int			bacc = 64 - __clzll(magAcc),
					brows = 64 - __clzll((unsigned long long) rows_seen);

		if (bb + bkc + bkd >= 63 || bacc + brows >= 63)
			atomicExch(P.status, CBGPU_ERR_OVERFLOW);
	}

	/* CTA reduction: warp shuffle, then one 128-bit shared add per warp, group and sum */
	const int	lane = threadIdx.x & 31;

#pragma unroll
	for (int g = 0; g < G; g++)
	{
#pragma unroll
		for (int s = 0; s <= SA_NSUM; s++)
		{
			long long	v = s < SA_NSUM ? acc.s[g][s < SA_NSUM ? s : 0] : (long long) acc.cnt[g];

			/* per-thread partials fit 63 bits (audited above); 32 of them need up to 68 bits, so
			 * reduce (lo, hi) pairs */
			unsigned long long lo = (unsigned long long) v;
			long long	hi = v < 0 ? -1 : 0;

#pragma unroll
			for (int o = 16; o; o >>= 1)
			{
				unsigned long long lo2 = __shfl_xor_sync(0xffffffffu, lo, o);
				long long	hi2 = __shfl_xor_sync(0xffffffffu, hi, o);
				unsigned long long nl = lo + lo2;

				hi = hi + hi2 + (nl < lo ? 1 : 0);
				lo = nl;
			}
			if (lane == 0 && (lo | (unsigned long long) hi))
			{
				unsigned long long old = atomicAdd(&red[g][s][0], lo);

				atomicAdd(&red[g][s][1], (unsigned long long) hi + (old + lo < old ? 1ull : 0ull));
			}
		}
	}
	__syncthreads();

	/* park the CTA's totals; k_small_commit folds them into the aggregate table */
	for (int i = threadIdx.x; i < G * (SA_NSUM + 1) * 2; i += blockDim.x)
		P.scratch[(size_t) blockIdx.x * G * (SA_NSUM + 1) * 2 + i] = (&red[0][0][0])[i];
	if (threadIdx.x < G)
		P.skeys[(size_t) blockIdx.x * G + threadIdx.x] = gkeys[threadIdx.x];
}

/* commit each (CTA, group) partial once: find-or-insert by the reference's group hash
 * (TupleHashTableHash_internal, executor/execGrouping.c:437-495), then exact 128-bit adds */
__global__ void
k_small_commit(const __grid_constant__ SmallAggParams P, int G, int nblocks)
{
	if (*P.retry)
		return;
	for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nblocks * G; e += gridDim.x * blockDim.x)
	{
		unsigned	key = P.skeys[e];
		const unsigned long long *r = P.scratch + (size_t) e * (SA_NSUM + 1) * 2;
		unsigned long long cnt = r[SA_NSUM * 2];

		if (key == 0 || cnt == 0)
			continue;
		int64_t		kv[2] = {(int64_t) (key & 0xff), (int64_t) ((key >> 8) & 0xff)};
		uint32_t	h = 0;

		for (int k = 0; k < P.nkeys; k++)
			h = pg_hash_combine(h, pg_hash_datum(P.keytype[k], kv[k], P.keydict[k]), false);
		h = pg_murmurhash32(h);
		int			slot = agg_find_or_insert(P.agg, h, kv, 0);

		if (slot < 0)
			continue;
		for (int a = 0; a < P.naccs; a++)
		{
			int			s = P.accsum[a];

			atomicAdd((unsigned long long *) (P.agg.n + (size_t) slot * P.agg.naccs + a), cnt);
			if (s >= 0)
				atomic_add128(P.agg.sum + ((size_t) slot * P.agg.naccs + a) * 2, r[s * 2], r[s * 2 + 1]);
		}
	}
}

/* ---------------------------------------------------------------------------------------------
 * matcher: does the descriptor have the small-group scan->agg shape?
 * --------------------------------------------------------------------------------------------- */
static bool
col_plain(const CbPipeline *p, int c, int width)
{
	return p->cols[c].src == 0 && p->cols[c].nulls == NULL && cb_type_w(p->cols[c].type) == width &&
		p->cols[c].type != CB_FLOAT8;
}

static int
try_small_agg(cbgpu_ctx *ctx, const CbPipeline *p, const PipeDev *d, bool *handled)
{
	static XProg x;
	SmallAggParams P;
	int			colB = -1,
				colC = -1,
				colD = -1,
				colA = -1;
	int64_t		k = 0,
				k2 = 0;
	bool		have_rev = false,
				have_chg = false;
	const CbpSink *s = &p->sink;

	*handled = false;
	if (s->kind != CBP_SINK_AGG || p->nprobes != 0 || p->drv_nsrc != 0 || s->nkeys > 2)
		return CBGPU_OK;
	if (!xm_decompile(p, &x) || x.depth < s->nkeys)
		return CBGPU_OK;
	memset(&P, 0, sizeof(P));
	/* quals: at most one CMP(int32 column, const) */
	if (x.nsections > 1)
		return CBGPU_OK;
	if (x.nsections == 1)
	{
		int			code,
					col;
		int64_t		v;

		if (x.sections[0].kind != 0 || !xm_is_cmp_const(&x, x.sections[0].node, &code, &col, &v) ||
			!col_plain(p, col, 4) || v < INT32_MIN || v > INT32_MAX)
			return CBGPU_OK;
		P.fcol = (const int32_t *) p->cols[col].data;
		P.fcode = code;
		P.fconst = (int32_t) v;
	}
	/* keys: one-byte columns */
	for (int i = 0; i < s->nkeys; i++)
	{
		int			col;

		if (!xm_is_load(&x, x.stack[i], &col) || !col_plain(p, col, 1))
			return CBGPU_OK;
		if (i == 0)
			P.key0 = (const uint8_t *) p->cols[col].data;
		else
			P.key1 = (const uint8_t *) p->cols[col].data;
		P.keytype[i] = s->keytype[i];
		P.keydict[i] = s->key_dict_hash[i];
	}
	/* first pass over the accumulators: find the product terms, they fix b, c, d */
	for (int a = 0; a < s->naccs; a++)
	{
		int			b,
					c,
					dd;
		int64_t		kk,
					kk2;

		if (s->accs[a].kind == CBP_ACC_COUNT && s->accs[a].arg < 0)
			continue;
		if (s->accs[a].kind != CBP_ACC_SUM_INT)
			return CBGPU_OK;
		int			node = x.stack[s->nkeys + s->accs[a].arg];

		if (xm_is_chg(&x, node, &b, &kk, &c, &kk2, &dd))
		{
			if ((have_rev || have_chg) && (b != colB || c != colC || kk != k))
				return CBGPU_OK;
			if (have_chg && (dd != colD || kk2 != k2))
				return CBGPU_OK;
			colB = b; colC = c; colD = dd; k = kk; k2 = kk2;
			have_chg = true;
		}
		else if (xm_is_rev(&x, node, &b, &kk, &c))
		{
			if ((have_rev || have_chg) && (b != colB || c != colC || kk != k))
				return CBGPU_OK;
			colB = b; colC = c; k = kk;
			have_rev = true;
		}
	}
	if (!have_rev && !have_chg)
		return CBGPU_OK;		/* plain-column-only aggregates go to the generic kernel for now */
	/* second pass: map every accumulator onto one of the six sums */
	P.naccs = s->naccs;
	for (int a = 0; a < s->naccs; a++)
	{
		int			b,
					c,
					dd,
					col;
		int64_t		kk,
					kk2;

		if (s->accs[a].kind == CBP_ACC_COUNT)
		{
			P.accsum[a] = -1;
			continue;
		}
		int			node = x.stack[s->nkeys + s->accs[a].arg];

		if (xm_is_chg(&x, node, &b, &kk, &c, &kk2, &dd))
			P.accsum[a] = 5;
		else if (xm_is_rev(&x, node, &b, &kk, &c))
			P.accsum[a] = 4;
		else if (xm_is_load(&x, node, &col))
		{
			if (col == colB)
				P.accsum[a] = 1;
			else if (col == colC)
				P.accsum[a] = 2;
			else if (col == colD)
				P.accsum[a] = 3;
			else if (colA < 0 || col == colA)
			{
				colA = col;
				P.accsum[a] = 0;
			}
			else if (colD < 0)
			{
				colD = col;		/* a second free column rides in d's slot (k2 stays 0: unused) */
				P.accsum[a] = 3;
			}
			else
				return CBGPU_OK;
		}
		else
			return CBGPU_OK;
	}
	if (!col_plain(p, colB, 8) || !col_plain(p, colC, 8) || (colA >= 0 && !col_plain(p, colA, 8)) ||
		(colD >= 0 && !col_plain(p, colD, 8)))
		return CBGPU_OK;
	P.nrows = p->nrows;
	P.colA = colA >= 0 ? (const long long *) p->cols[colA].data : NULL;
	P.colB = (const long long *) p->cols[colB].data;
	P.colC = (const long long *) p->cols[colC].data;
	P.colD = colD >= 0 ? (const long long *) p->cols[colD].data : NULL;
	P.k = k;
	P.k2 = k2;
	P.visimap = p->visimap;
	P.agg = d->sink.agg;
	P.nkeys = s->nkeys;
	P.status = ctx->d_status;

	P.want_chg = have_chg;

	/* retry flag: more distinct groups in one CTA than register slots -> wider kernel, then generic */
	int		   *d_retry;
	int			h_retry = 0;
	int64_t		ntiles = (p->nrows + 2 * SA_THREADS - 1) / (2 * SA_THREADS);
	int			blocks = ctx->sm_count * 2;

	if (blocks > ntiles)
		blocks = (int) ntiles;
	CB_CUDA(ctx, cudaMallocAsync(&d_retry, sizeof(int), ctx->stream));
	CB_CUDA(ctx, cudaMallocAsync(&P.scratch, (size_t) blocks * 8 * (SA_NSUM + 1) * 2 * sizeof(unsigned long long), ctx->stream));
	CB_CUDA(ctx, cudaMallocAsync(&P.skeys, (size_t) blocks * 8 * sizeof(unsigned), ctx->stream));
	P.retry = d_retry;
	for (int attempt = 0; attempt < 2; attempt++)
	{
		int			G = attempt == 0 ? 4 : 8;

		CB_CUDA(ctx, cudaMemsetAsync(d_retry, 0, sizeof(int), ctx->stream));
		CB_CUDA(ctx, cudaEventRecord(ctx->ev_k0, ctx->stream));
		if (attempt == 0)
			k_scan_agg_small<4><<<blocks, SA_THREADS, 0, ctx->stream>>>(P);
		else
			k_scan_agg_small<8><<<blocks, SA_THREADS, 0, ctx->stream>>>(P);
		CB_LAUNCHED(ctx, "k_scan_agg_small");
		CB_CUDA(ctx, cudaEventRecord(ctx->ev_k1, ctx->stream));
		ctx->kernel_timed = true;
		ctx->last_kernel_name = attempt == 0 ? "k_scan_agg_small<4>" : "k_scan_agg_small<8>";
		k_small_commit<<<1, 256, 0, ctx->stream>>>(P, G, blocks);
		CB_LAUNCHED(ctx, "k_small_commit");
		CB_CUDA(ctx, cudaMemcpyAsync(&h_retry, d_retry, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
		CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		if (!h_retry)
		{
			*handled = true;
			break;
		}
	}
	CB_CUDA(ctx, cudaFreeAsync(d_retry, ctx->stream));
	CB_CUDA(ctx, cudaFreeAsync(P.scratch, ctx->stream));
	CB_CUDA(ctx, cudaFreeAsync(P.skeys, ctx->stream));
	return CBGPU_OK;
}

int
cb_try_specialised(cbgpu_ctx *ctx, const CbPipeline *p, const PipeDev *d, bool *handled)
{
	return try_small_agg(ctx, p, d, handled);
}
