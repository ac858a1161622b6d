/*
 * agg.cu - aggregate hash tables (allocation, read-back, conversion to a relation) and device top-N.
 *
 * The table stands in for the reference's TupleHashTable of AggStatePerGroupData
 * (backend/executor/execGrouping.c:317, backend/executor/nodeAgg.c:2220-2319).  Per group it holds
 * the grouping keys and, per accumulator, (N, 128-bit exact sum) - the same content as the
 * reference's Int128AggState / NumericAggState{N, sumX} (backend/utils/adt/numeric.c:5340, 4602)
 * or float8 {N, Sx} - so partial states can be redistributed and combined (two-stage aggregation,
 * nodes/nodes.h:977-1000 AggSplit) and finalised exactly on the host.
 */
#include "common.cuh"

#include <stdlib.h>

static int64_t
next_pow2(int64_t v)
{
	int64_t		p = 1;

	while (p < v)
		p <<= 1;
	return p;
}

struct AggKinds
{
	int32_t		k[CBP_MAX_AGGS];
};

__global__ void
k_agg_init(AggDev t, const AggKinds kk)
{
	const int32_t *kinds = kk.k;
	size_t		cap = (size_t) t.mask + 1;
	size_t		i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	size_t		stride = (size_t) gridDim.x * blockDim.x;

	for (; i < cap; i += stride)
	{
		t.state[i] = 0;
		for (int a = 0; a < t.naccs; a++)
		{
			unsigned long long lo = 0;

			if (kinds[a] == CBP_ACC_MIN || kinds[a] == CBP_ACC_MERGE_MIN)
				lo = 0x7fffffffffffffffull;
			else if (kinds[a] == CBP_ACC_MAX || kinds[a] == CBP_ACC_MERGE_MAX)
				lo = 0x8000000000000000ull;
			t.n[i * t.naccs + a] = 0;
			t.sum[(i * t.naccs + a) * 2] = lo;
			t.sum[(i * t.naccs + a) * 2 + 1] = 0;
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0)
	{
		*t.ngroups = 0;
		*t.full = 0;
	}
}

extern "C" int
cbgpu_agg_reset(cbgpu_aggtable *t)
{
	cbgpu_ctx  *ctx = t->ctx;
	AggKinds	kk;

	t->counted = false;
	t->snap_valid = false;
	memcpy(kk.k, t->kinds, sizeof(kk.k));
	int			blocks = (int) ((t->capacity + 255) / 256);

	if (blocks > ctx->sm_count * 8)
		blocks = ctx->sm_count * 8;
	k_agg_init<<<blocks, 256, 0, ctx->stream>>>(t->d, kk);
	CB_LAUNCHED(ctx, "k_agg_init");
	return CBGPU_OK;
}

extern "C" int
cbgpu_agg_create(cbgpu_ctx *ctx, int32_t nkeys, int32_t naccs, const int32_t *acc_kinds, int64_t capacity_groups,
				 cbgpu_aggtable **out)
{
	cbgpu_aggtable *t;
	int64_t		cap;

	if (nkeys < 0 || nkeys > CBP_MAX_KEYS || naccs < 0 || naccs > CBP_MAX_AGGS)
		return cb_fail(ctx, CBGPU_ERR_UNSUPPORTED, "aggregate with %s%lld keys/accumulators is beyond the GPU path's limits", "", (long long) (nkeys * 100 + naccs));
	t = (cbgpu_aggtable *) calloc(1, sizeof(cbgpu_aggtable));
	if (!t)
		return CBGPU_ERR_NOMEM;
	cap = next_pow2((capacity_groups < 16 ? 16 : capacity_groups) * 2);
	if (cap > (1ll << 31))
	{
		free(t);
		return cb_fail(ctx, CBGPU_ERR_UNSUPPORTED, "aggregate table of %s%lld slots exceeds the GPU path's limit", "", cap);
	}
	t->ctx = ctx;
	t->capacity = cap;
	t->d.mask = (uint32_t) (cap - 1);
	t->d.nkeys = nkeys;
	t->d.naccs = naccs;
	for (int a = 0; a < naccs; a++)
		t->kinds[a] = acc_kinds ? acc_kinds[a] : CBP_ACC_SUM_INT;
	CB_CUDA(ctx, cudaSetDevice(ctx->device));
	{
		/* one allocation for all the arrays (a table is created per query: seven pool calls were a measurable part of a
		 * 4 ms step) */
		size_t		off[8];
		size_t		sz[7] = {(size_t) cap * sizeof(int32_t), (size_t) cap * sizeof(uint32_t), (size_t) cap * sizeof(int64_t) * (nkeys ? nkeys : 1),
							 (size_t) cap * sizeof(uint32_t), (size_t) cap * sizeof(int64_t) * (naccs ? naccs : 1),
							 (size_t) cap * 2 * sizeof(unsigned long long) * (naccs ? naccs : 1), sizeof(int32_t) * 4};
		char	   *base;

		off[0] = 0;
		for (int i = 0; i < 7; i++)
			off[i + 1] = off[i] + ((sz[i] + 255) & ~(size_t) 255);
		CB_CUDA(ctx, cudaMallocAsync(&base, off[7], ctx->stream));
		t->base = base;
		t->d.state = (int32_t *) (base + off[0]);
		t->d.hash = (uint32_t *) (base + off[1]);
		t->d.keys = (int64_t *) (base + off[2]);
		t->d.keynull = (uint32_t *) (base + off[3]);
		t->d.n = (int64_t *) (base + off[4]);
		t->d.sum = (unsigned long long *) (base + off[5]);
		t->d.ngroups = (int32_t *) (base + off[6]);
	}
	t->d.full = t->d.ngroups + 1;
	*out = t;
	return cbgpu_agg_reset(t);
}

/* device bytes one slot of an aggregate table takes (two slots per group are provisioned: load factor <= 0.5) */
extern "C" int64_t
cbgpu_agg_slot_bytes(int32_t nkeys, int32_t naccs)
{
	return 4 + 4 + 4 + 8 * (int64_t) (nkeys ? nkeys : 1) + (8 + 16) * (int64_t) (naccs ? naccs : 1);
}

extern "C" int
cbgpu_agg_set_partition(cbgpu_aggtable *t, int32_t npart, int32_t part)
{
	if (npart < 1 || npart > 65536 || (npart & (npart - 1)) != 0 || part < 0 || part >= npart)
		return cb_fail(t->ctx, CBGPU_ERR_INVALID, "aggregate partition %s%lld out of range", "", part);
	t->d.npart = npart;
	t->d.part_id = part;
	t->d.part_shift = 32;
	for (int b = npart; b > 1; b >>= 1)
		t->d.part_shift--;
	return CBGPU_OK;
}

extern "C" void
cbgpu_agg_free(cbgpu_aggtable *t)
{
	if (!t)
		return;
	cudaSetDevice(t->ctx->device);
	cudaFreeAsync(t->base, t->ctx->stream);
	free(t->snap);
	free(t);
}

/* number of published groups (TupleHashTable's `members`), and whether any group key is NULL */
__global__ void
k_agg_count(AggDev t, int *anynull)
{
	size_t		cap = (size_t) t.mask + 1;
	size_t		i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	size_t		stride = (size_t) gridDim.x * blockDim.x;
	int			c = 0;
	int			an = 0;

	for (; i < cap; i += stride)
		if (t.state[i] == 2)
		{
			c++;
			an |= t.keynull[i] != 0;
		}
	for (int o = 16; o > 0; o >>= 1)
	{
		c += __shfl_xor_sync(0xffffffffu, c, o);
		an |= __shfl_xor_sync(0xffffffffu, an, o);
	}
	if ((threadIdx.x & 31) == 0)
	{
		if (c)
			atomicAdd(t.ngroups, c);
		if (an)
			*anynull = 1;
	}
}

/* small tables: count, flags and the groups themselves in one kernel and one round trip */
__global__ void __launch_bounds__(256)
k_agg_snapshot(AggDev t, AggSnap *out)
{
	if (threadIdx.x == 0)
		out->retry = out->audit = 0;
	agg_snapshot_block(t, out);
}

/* the snapshot the stream has just been synchronised on describes table t */
void
cb_agg_adopt_snapshot(cbgpu_aggtable *t, const AggSnap *snap)
{
	t->snap_valid = false;
	if (snap->ngroups < 0)
		return;
	t->ngroups = snap->ngroups;
	t->anynull = snap->anynull;
	t->counted = !snap->full;
	if (snap->full || snap->ngroups > AGG_SNAP_MAXG)
		return;
	if (!t->snap)
		t->snap = (AggSnap *) malloc(sizeof(AggSnap));
	if (!t->snap)
		return;
	memcpy(t->snap, snap, sizeof(AggSnap));
	t->snap_valid = true;
}

extern "C" int
cbgpu_agg_ngroups(cbgpu_aggtable *t, int64_t *ngroups)
{
	cbgpu_ctx  *ctx = t->ctx;
	int32_t		h[3];

	if (!t->counted && t->capacity <= AGG_SNAP_MAXCAP && ctx->agg_snap)
	{
		ctx->agg_snap->ngroups = -1;
		k_agg_snapshot<<<1, 256, 0, ctx->stream>>>(t->d, ctx->agg_snap);
		CB_LAUNCHED(ctx, "k_agg_snapshot");
		CB_CUDA(ctx, CB_STATUS_RIDE(ctx));
		CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		CB_STATUS_FETCHED(ctx);
		if (ctx->agg_snap->full)
			return cb_fail(ctx, CBGPU_ERR_NOMEM, "aggregate hash table overflow (%s capacity %lld slots)", "", t->capacity);
		cb_agg_adopt_snapshot(t, ctx->agg_snap);
	}
	if (!t->counted)
	{
		/* groups are counted here, once per fill, rather than with one same-address atomic per new
		 * group; d.ngroups = {count, table-full flag, any-NULL-key flag} */
		CB_CUDA(ctx, cudaMemsetAsync(t->d.ngroups, 0, sizeof(int32_t), ctx->stream));
		CB_CUDA(ctx, cudaMemsetAsync(t->d.ngroups + 2, 0, sizeof(int32_t), ctx->stream));
		{
			int			blocks = (int) ((t->capacity + 1023) / 1024);

			if (blocks > ctx->sm_count * 8)
				blocks = ctx->sm_count * 8;
			k_agg_count<<<blocks, 256, 0, ctx->stream>>>(t->d, t->d.ngroups + 2);
			CB_LAUNCHED(ctx, "k_agg_count");
		}
		CB_CUDA(ctx, cudaMemcpyAsync(h, t->d.ngroups, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
		CB_CUDA(ctx, CB_STATUS_RIDE(ctx));
		CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		CB_STATUS_FETCHED(ctx);
		if (h[1])
			return cb_fail(ctx, CBGPU_ERR_NOMEM, "aggregate hash table overflow (%s capacity %lld slots)", "", t->capacity);
		t->ngroups = h[0];
		t->anynull = h[2];
		t->counted = true;
	}
	*ngroups = t->ngroups;
	return CBGPU_OK;
}

/* the table is about to be written (a pipeline's AGG sink, a reset): forget the cached count */
void
cb_agg_touch(cbgpu_aggtable *t)
{
	t->counted = false;
	t->snap_valid = false;
}

/* compaction: slots in `ready` state -> dense rows (agg_retrieve_hash_table's table walk,
 * backend/executor/nodeAgg.c:3014) */
struct AggOut
{
	int64_t    *keys;
	uint32_t   *keynull;
	int64_t    *n;
	int64_t    *lo;
	int64_t    *hi;
	int32_t    *counter;
	int64_t		maxgroups;
};

__global__ void
k_agg_compact(AggDev t, AggOut o)
{
	size_t		cap = (size_t) t.mask + 1;
	size_t		i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	size_t		stride = (size_t) gridDim.x * blockDim.x;

	for (; i < cap; i += stride)
	{
		if (t.state[i] != 2)
			continue;
		int			g = atomicAdd(o.counter, 1);

		if (g >= o.maxgroups)
			continue;
		for (int k = 0; k < t.nkeys; k++)
			o.keys[(size_t) g * t.nkeys + k] = t.keys[i * t.nkeys + k];
		o.keynull[g] = t.keynull[i];
		for (int a = 0; a < t.naccs; a++)
		{
			o.n[(size_t) g * t.naccs + a] = t.n[i * t.naccs + a];
			o.lo[(size_t) g * t.naccs + a] = (int64_t) t.sum[(i * t.naccs + a) * 2];
			o.hi[(size_t) g * t.naccs + a] = (int64_t) t.sum[(i * t.naccs + a) * 2 + 1];
		}
	}
}

extern "C" int
cbgpu_agg_read(cbgpu_aggtable *t, int64_t maxgroups, int64_t *keys, uint32_t *keynull, int64_t *n, int64_t *sum_lo,
			   int64_t *sum_hi, int64_t *ngroups)
{
	cbgpu_ctx  *ctx = t->ctx;
	int64_t		ng;
	int			rc = cbgpu_agg_ngroups(t, &ng);
	AggOut		o;
	int			nk = t->d.nkeys ? t->d.nkeys : 1;
	int			na = t->d.naccs ? t->d.naccs : 1;

	if (rc)
		return rc;
	*ngroups = ng;
	if (ng > maxgroups)
		return cb_fail(ctx, CBGPU_ERR_INVALID, "cbgpu_agg_read: %s%lld groups exceed the caller's buffer", "", ng);
	if (ng == 0)
		return CBGPU_OK;
	if (t->snap_valid && t->snap->ngroups == ng)
	{
		/* the groups came with the count */
		for (int64_t g = 0; g < ng; g++)
		{
			keynull[g] = t->snap->keynull[g];
			for (int k = 0; k < t->d.nkeys; k++)
				keys[g * t->d.nkeys + k] = t->snap->keys[g][k];
			for (int a = 0; a < t->d.naccs; a++)
			{
				n[g * t->d.naccs + a] = t->snap->n[g][a];
				sum_lo[g * t->d.naccs + a] = t->snap->lo[g][a];
				sum_hi[g * t->d.naccs + a] = t->snap->hi[g][a];
			}
		}
		return CBGPU_OK;
	}
	o.maxgroups = ng;
	CB_CUDA(ctx, cudaMallocAsync(&o.keys, sizeof(int64_t) * ng * nk, ctx->stream));
	CB_CUDA(ctx, cudaMallocAsync(&o.keynull, sizeof(uint32_t) * ng, ctx->stream));
	CB_CUDA(ctx, cudaMallocAsync(&o.n, sizeof(int64_t) * ng * na, ctx->stream));
	CB_CUDA(ctx, cudaMallocAsync(&o.lo, sizeof(int64_t) * ng * na, ctx->stream));
	CB_CUDA(ctx, cudaMallocAsync(&o.hi, sizeof(int64_t) * ng * na, ctx->stream));
	CB_CUDA(ctx, cudaMallocAsync(&o.counter, sizeof(int32_t), ctx->stream));
	CB_CUDA(ctx, cudaMemsetAsync(o.counter, 0, sizeof(int32_t), ctx->stream));
	int			blocks = (int) ((t->capacity + 255) / 256);

	if (blocks > ctx->sm_count * 8)
		blocks = ctx->sm_count * 8;
	k_agg_compact<<<blocks, 256, 0, ctx->stream>>>(t->d, o);
	CB_LAUNCHED(ctx, "k_agg_compact");
	if (t->d.nkeys)
		CB_CUDA(ctx, cudaMemcpyAsync(keys, o.keys, sizeof(int64_t) * ng * t->d.nkeys, cudaMemcpyDeviceToHost, ctx->stream));
	CB_CUDA(ctx, cudaMemcpyAsync(keynull, o.keynull, sizeof(uint32_t) * ng, cudaMemcpyDeviceToHost, ctx->stream));
	if (t->d.naccs)
	{
		CB_CUDA(ctx, cudaMemcpyAsync(n, o.n, sizeof(int64_t) * ng * t->d.naccs, cudaMemcpyDeviceToHost, ctx->stream));
		CB_CUDA(ctx, cudaMemcpyAsync(sum_lo, o.lo, sizeof(int64_t) * ng * t->d.naccs, cudaMemcpyDeviceToHost, ctx->stream));
		CB_CUDA(ctx, cudaMemcpyAsync(sum_hi, o.hi, sizeof(int64_t) * ng * t->d.naccs, cudaMemcpyDeviceToHost, ctx->stream));
	}
	CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	cudaFreeAsync(o.keys, ctx->stream);
	cudaFreeAsync(o.keynull, ctx->stream);
	cudaFreeAsync(o.n, ctx->stream);
	cudaFreeAsync(o.lo, ctx->stream);
	cudaFreeAsync(o.hi, ctx->stream);
	cudaFreeAsync(o.counter, ctx->stream);
	return CBGPU_OK;
}

/* groups -> relation: key columns (typed), then per accumulator N, sum.lo, sum.hi (int8 each) */
struct AggRelOut
{
	void	   *key[CBP_MAX_KEYS];
	uint8_t    *keynulls[CBP_MAX_KEYS];
	int32_t		keytype[CBP_MAX_KEYS];
	int64_t    *acc[CBP_MAX_AGGS * 3];
	int32_t    *counter;
};

__global__ void
k_agg_to_rel(AggDev t, AggRelOut o)
{
	size_t		cap = (size_t) t.mask + 1;
	size_t		i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	size_t		stride = (size_t) gridDim.x * blockDim.x;

	for (; i < cap; i += stride)
	{
		if (t.state[i] != 2)
			continue;
		int			g = atomicAdd(o.counter, 1);
		uint32_t	nm = t.keynull[i];

		for (int k = 0; k < t.nkeys; k++)
		{
			int64_t		v = t.keys[i * t.nkeys + k];

			switch (o.keytype[k])
			{
				case CB_INT4: case CB_DATE: case CB_DICT32:
					((int32_t *) o.key[k])[g] = (int32_t) v;
					break;
				case CB_INT8: case CB_NUMERIC: case CB_FLOAT8:
					((int64_t *) o.key[k])[g] = v;
					break;
				default:
					((uint8_t *) o.key[k])[g] = (uint8_t) v;
			}
			if (o.keynulls[k])
				o.keynulls[k][g] = (nm >> k) & 1;
		}
		for (int a = 0; a < t.naccs; a++)
		{
			o.acc[a * 3][g] = t.n[i * t.naccs + a];
			o.acc[a * 3 + 1][g] = (int64_t) t.sum[(i * t.naccs + a) * 2];
			o.acc[a * 3 + 2][g] = (int64_t) t.sum[(i * t.naccs + a) * 2 + 1];
		}
	}
}

__global__ void
k_agg_anynull(AggDev t, int *flag)
{
	size_t		cap = (size_t) t.mask + 1;
	size_t		i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	size_t		stride = (size_t) gridDim.x * blockDim.x;

	for (; i < cap; i += stride)
		if (t.state[i] == 2 && t.keynull[i])
			*flag = 1;
}

extern "C" int
cbgpu_agg_to_rel(cbgpu_aggtable *t, const int32_t *keytypes, cbgpu_rel **out)
{
	cbgpu_ctx  *ctx = t->ctx;
	int64_t		ng;
	int			rc = cbgpu_agg_ngroups(t, &ng);
	int32_t		types[CBP_MAX_KEYS + CBP_MAX_AGGS * 3];
	int			ncols = t->d.nkeys + t->d.naccs * 3;
	cbgpu_rel  *rel;
	AggRelOut	o;
	int			h_flag = 0;

	if (rc)
		return rc;
	for (int k = 0; k < t->d.nkeys; k++)
		types[k] = keytypes[k];
	for (int a = 0; a < t->d.naccs * 3; a++)
		types[t->d.nkeys + a] = CB_INT8;
	rc = cbgpu_rel_create(ctx, ng, ncols, types, NULL, &rel);
	if (rc)
		return rc;
	memset(&o, 0, sizeof(o));
	int			blocks = (int) ((t->capacity + 255) / 256);

	if (blocks > ctx->sm_count * 8)
		blocks = ctx->sm_count * 8;
	/* NULL group keys need null byte-maps on the key columns (flag gathered with the group count) */
	h_flag = t->anynull;
	for (int k = 0; k < t->d.nkeys; k++)
	{
		o.key[k] = rel->data[k];
		o.keytype[k] = keytypes[k];
		if (h_flag && ng > 0)
		{
			CB_CUDA(ctx, cudaMallocAsync(&rel->nulls[k], (size_t) ng, ctx->stream));
			o.keynulls[k] = rel->nulls[k];
		}
	}
	for (int a = 0; a < t->d.naccs * 3; a++)
		o.acc[a] = (int64_t *) rel->data[t->d.nkeys + a];
	CB_CUDA(ctx, cudaMallocAsync(&o.counter, sizeof(int32_t), ctx->stream));
	CB_CUDA(ctx, cudaMemsetAsync(o.counter, 0, sizeof(int32_t), ctx->stream));
	if (ng > 0)
	{
		k_agg_to_rel<<<blocks, 256, 0, ctx->stream>>>(t->d, o);
		CB_LAUNCHED(ctx, "k_agg_to_rel");
	}
	CB_CUDA(ctx, cudaFreeAsync(o.counter, ctx->stream));
	*out = rel;
	return CBGPU_OK;
}

/* ---------------------------------------------------------------------------------------------
 * top-N: ORDER BY keys LIMIT n over a relation (bounded sort above the Agg)
 * --------------------------------------------------------------------------------------------- */
#define TOPN_MAXK 64
#define TOPN_MAXKEYS 4

struct TopnParams
{
	const void *key[TOPN_MAXKEYS];
	int32_t		keytype[TOPN_MAXKEYS];
	int32_t		desc[TOPN_MAXKEYS];
	int32_t		uns[TOPN_MAXKEYS];
	int32_t		nkeys;
	int32_t		k;
	const uint32_t *in_idx;		/* NULL = identity over nrows                                         */
	int64_t		nin;
	uint32_t   *out_idx;		/* [nthreads * k], padded with 0xFFFFFFFF                             */
};

/* true when row a sorts strictly before row b */
__device__ __forceinline__ bool
topn_before(const TopnParams &p, const int64_t *ka, uint32_t ra, const int64_t *kb, uint32_t rb)
{
	for (int i = 0; i < p.nkeys; i++)
	{
		int64_t		x = ka[i],
					y = kb[i];

		if (x == y)
			continue;
		bool		lt = p.uns[i] ? ((uint64_t) x < (uint64_t) y) : (x < y);

		return p.desc[i] ? !lt : lt;
	}
	return ra < rb;
}

__global__ void
k_topn(TopnParams p)
{
	uint32_t	best_row[TOPN_MAXK];
	int64_t		best_key[TOPN_MAXK][TOPN_MAXKEYS];
	int			nbest = 0;
	int64_t		tid = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	int64_t		stride = (int64_t) gridDim.x * blockDim.x;

	for (int64_t i = tid; i < p.nin; i += stride)
	{
		uint32_t	row = p.in_idx ? p.in_idx[i] : (uint32_t) i;
		int64_t		key[TOPN_MAXKEYS];

		if (row == 0xFFFFFFFFu)
			continue;
		for (int k = 0; k < p.nkeys; k++)
			key[k] = cb_load_widen(p.key[k], p.keytype[k], row);
		if (nbest == p.k && !topn_before(p, key, row, best_key[nbest - 1], best_row[nbest - 1]))
			continue;
		/* insertion into the sorted local list */
		int			pos = nbest < p.k ? nbest : p.k - 1;

		while (pos > 0 && topn_before(p, key, row, best_key[pos - 1], best_row[pos - 1]))
		{
			best_row[pos] = best_row[pos - 1];
			for (int k = 0; k < p.nkeys; k++)
				best_key[pos][k] = best_key[pos - 1][k];
			pos--;
		}
		best_row[pos] = row;
		for (int k = 0; k < p.nkeys; k++)
			best_key[pos][k] = key[k];
		if (nbest < p.k)
			nbest++;
	}
	for (int j = 0; j < p.k; j++)
		p.out_idx[tid * p.k + j] = j < nbest ? best_row[j] : 0xFFFFFFFFu;
}

/* --- threshold selection: the usual case (n >> k) costs two streaming passes over ONE key column ---
 * A: per-CTA minimum of the leading non-constant sort key, mapped so that smaller = sorts first;
 * K: the k-th smallest CTA minimum T bounds the answer (k rows - those minima - are <= T);
 * B: rows with key <= T become candidates (a few dozen when the key is not heavily tied);
 * C: one CTA ranks the candidates with the full comparator. */
#define TOPN_BLOCKS 1024
#define TOPN_CAND 4096

__device__ __forceinline__ unsigned long long
topn_ukey(const TopnParams &p, int i, uint32_t row)
{
	unsigned long long u = (unsigned long long) cb_load_widen(p.key[i], p.keytype[i], row);

	if (!p.uns[i])
		u ^= 0x8000000000000000ull;
	return p.desc[i] ? ~u : u;
}

__global__ void __launch_bounds__(256)
k_topn_min(TopnParams p, int keyi, unsigned long long *bmin, unsigned long long *bmax)
{
	__shared__ unsigned long long smin[8], smax[8];
	unsigned long long lo = ~0ull,
				hi = 0;

	for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < p.nin; i += (int64_t) gridDim.x * blockDim.x)
	{
		const unsigned long long u = topn_ukey(p, keyi, (uint32_t) i);

		lo = u < lo ? u : lo;
		hi = u > hi ? u : hi;
	}
	for (int o = 16; o > 0; o >>= 1)
	{
		const unsigned long long a = __shfl_xor_sync(0xffffffffu, lo, o);
		const unsigned long long b = __shfl_xor_sync(0xffffffffu, hi, o);

		lo = a < lo ? a : lo;
		hi = b > hi ? b : hi;
	}
	if ((threadIdx.x & 31) == 0)
	{
		smin[threadIdx.x >> 5] = lo;
		smax[threadIdx.x >> 5] = hi;
	}
	__syncthreads();
	if (threadIdx.x == 0)
	{
		for (int w = 1; w < 8; w++)
		{
			lo = smin[w] < lo ? smin[w] : lo;
			hi = smax[w] > hi ? smax[w] : hi;
		}
		bmin[blockIdx.x] = lo;
		bmax[blockIdx.x] = hi;
	}
}

/* out[0] = k-th smallest block minimum (or the maximum when there are fewer than k blocks),
 * out[1] = global minimum, out[2] = global maximum */
__global__ void __launch_bounds__(TOPN_BLOCKS)
k_topn_kth(const unsigned long long *bmin, const unsigned long long *bmax, int nblocks, int k, unsigned long long *out)
{
	__shared__ unsigned long long v[TOPN_BLOCKS];
	__shared__ unsigned long long gmax;
	const int	t = threadIdx.x;

	v[t] = t < nblocks ? bmin[t] : ~0ull;
	if (t == 0)
	{
		unsigned long long m = 0;

		for (int i = 0; i < nblocks; i++)
			m = bmax[i] > m ? bmax[i] : m;
		gmax = m;
	}
	__syncthreads();
	for (int size = 2; size <= TOPN_BLOCKS; size <<= 1)
		for (int stride = size >> 1; stride > 0; stride >>= 1)
		{
			const int	partner = t ^ stride;

			if (partner > t)
			{
				const bool	up = (t & size) == 0;
				const unsigned long long a = v[t],
							b = v[partner];

				if ((a > b) == up)
				{
					v[t] = b;
					v[partner] = a;
				}
			}
			__syncthreads();
		}
	if (t == 0)
	{
		out[0] = nblocks >= k ? v[k - 1] : gmax;
		out[1] = v[0];
		out[2] = gmax;
	}
}

__global__ void __launch_bounds__(256)
k_topn_filter(TopnParams p, int keyi, const unsigned long long *thr, uint32_t *cand, int *ncand)
{
	const unsigned long long T = thr[0];

	for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < p.nin; i += (int64_t) gridDim.x * blockDim.x)
		if (topn_ukey(p, keyi, (uint32_t) i) <= T)
		{
			const int	pos = atomicAdd(ncand, 1);

			if (pos < TOPN_CAND)
				cand[pos] = (uint32_t) i;
		}
}

/* rank every candidate among the candidates with the full comparator; ranks < k are the answer */
__global__ void __launch_bounds__(1024)
k_topn_final(TopnParams p, const uint32_t *cand, const int *ncand, int64_t nall, uint32_t *out)
{
	const int	m = cand ? (*ncand < TOPN_CAND ? *ncand : TOPN_CAND) : (int) nall;

	for (int j = threadIdx.x; j < p.k; j += blockDim.x)
		out[j] = 0xFFFFFFFFu;
	__syncthreads();
	for (int i = threadIdx.x; i < m; i += blockDim.x)
	{
		const uint32_t ri = cand ? cand[i] : (uint32_t) i;
		int64_t		ki[TOPN_MAXKEYS];
		int			rank = 0;

		for (int k = 0; k < p.nkeys; k++)
			ki[k] = cb_load_widen(p.key[k], p.keytype[k], ri);
		for (int j = 0; j < m && rank < p.k; j++)
		{
			const uint32_t rj = cand ? cand[j] : (uint32_t) j;
			int64_t		kj[TOPN_MAXKEYS];

			if (j == i)
				continue;
			for (int k = 0; k < p.nkeys; k++)
				kj[k] = cb_load_widen(p.key[k], p.keytype[k], rj);
			rank += topn_before(p, kj, rj, ki, ri);
		}
		if (rank < p.k)
			out[rank] = ri;
	}
}

/* ---------------------------------------------------------------------------------------------
 * merge receive: a Gather Motion whose senders' streams are sorted (Motion.sendSorted; execMotionSortedReceiver,
 * nodeMotion.c:433, merges them with a binary heap under CdbMergeComparator :1010).  The gathered relation holds the
 * streams one after another, so it is a sequence of sorted runs: a row's place in the merged order is its place in its
 * own run plus, for every other run, the number of rows there that go before it (binary search; equal keys: the earlier
 * run first) - every row finds its place independently, no heap, no passes.
 * --------------------------------------------------------------------------------------------- */
#define MERGE_MAXRUNS 64

struct MergeParams
{
	const void *key[TOPN_MAXKEYS];
	const uint8_t *nulls[TOPN_MAXKEYS];
	int32_t		keytype[TOPN_MAXKEYS];
	int32_t		desc[TOPN_MAXKEYS];
	int32_t		uns[TOPN_MAXKEYS];
	int32_t		nkeys;
	int32_t		nruns;
	int64_t		n;
	int64_t		start[MERGE_MAXRUNS + 1];
	int32_t    *nfound;			/* run starts found (k_merge_find_runs)                               */
	unsigned long long *found;
	uint32_t   *order;
};

/* < 0: row a sorts before row b; 0: equal keys.  NULLS LAST ascending, NULLS FIRST descending (the defaults) */
__device__ __forceinline__ int
merge_cmp(const MergeParams &p, int64_t a, int64_t b)
{
	for (int i = 0; i < p.nkeys; i++)
	{
		const bool	an = p.nulls[i] && p.nulls[i][a];
		const bool	bn = p.nulls[i] && p.nulls[i][b];
		int			c;

		if (an || bn)
			c = (an && bn) ? 0 : (an ? 1 : -1);
		else
		{
			const int64_t x = cb_load_widen(p.key[i], p.keytype[i], (uint32_t) a);
			const int64_t y = cb_load_widen(p.key[i], p.keytype[i], (uint32_t) b);

			if (x == y)
				c = 0;
			else if (p.uns[i])
				c = (uint64_t) x < (uint64_t) y ? -1 : 1;
			else
				c = x < y ? -1 : 1;
		}
		if (p.desc[i])
			c = -c;
		if (c)
			return c;
	}
	return 0;
}

/* a run starts wherever a row sorts before its predecessor */
__global__ void
k_merge_find_runs(MergeParams p)
{
	int64_t		i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x + 1;
	const int64_t stride = (int64_t) gridDim.x * blockDim.x;

	for (; i < p.n; i += stride)
		if (merge_cmp(p, i, i - 1) < 0)
		{
			const int	at = atomicAdd(p.nfound, 1);

			if (at < MERGE_MAXRUNS)
				p.found[at] = (unsigned long long) i;
		}
}

__global__ void
k_merge_place(MergeParams p)
{
	int64_t		i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t stride = (int64_t) gridDim.x * blockDim.x;

	for (; i < p.n; i += stride)
	{
		int			r = 0;
		int64_t		pos;

		while (r + 1 < p.nruns && p.start[r + 1] <= i)
			r++;
		pos = i - p.start[r];
		for (int q = 0; q < p.nruns; q++)
		{
			int64_t		lo = p.start[q],
						hi = p.start[q + 1];

			if (q == r)
				continue;
			/* rows of run q that go before row i: those that sort before it, and - in an earlier run - its equals too */
			while (lo < hi)
			{
				const int64_t mid = lo + (hi - lo) / 2;
				const int	c = merge_cmp(p, mid, i);

				if (c < 0 || (c == 0 && q < r))
					lo = mid + 1;
				else
					hi = mid;
			}
			pos += lo - p.start[q];
		}
		p.order[pos] = (uint32_t) i;
	}
}

extern "C" int
cbgpu_merge_sorted_runs(cbgpu_ctx *ctx, cbgpu_rel *rel, const int32_t *keycols, const int32_t *descending, const int32_t *unsigned_cmp,
						int32_t nkeys, int32_t max_runs, uint32_t **order_dev, int32_t *nruns_out)
{
	MergeParams p;
	int32_t		nfound = 0;
	unsigned long long found[MERGE_MAXRUNS];
	int			blocks;

	*order_dev = NULL;
	if (nruns_out)
		*nruns_out = rel->nrows > 0;
	if (nkeys < 1 || nkeys > TOPN_MAXKEYS || max_runs < 1 || max_runs > MERGE_MAXRUNS)
		return cb_fail(ctx, CBGPU_ERR_UNSUPPORTED, "merge receive supports 1..4 sort keys and up to 64 senders (%s got %lld keys)", "", nkeys);
	memset(&p, 0, sizeof(p));
	for (int i = 0; i < nkeys; i++)
	{
		if (keycols[i] < 0 || keycols[i] >= rel->ncols)
			return cb_fail(ctx, CBGPU_ERR_INVALID, "cbgpu_merge_sorted_runs: bad key column%s %lld", "", keycols[i]);
		p.key[i] = rel->data[keycols[i]];
		p.nulls[i] = rel->nulls[keycols[i]];
		p.keytype[i] = rel->types[keycols[i]];
		p.desc[i] = descending[i];
		p.uns[i] = unsigned_cmp ? unsigned_cmp[i] : 0;
		if (p.keytype[i] == CB_FLOAT8)
			return cb_fail(ctx, CBGPU_ERR_UNSUPPORTED, "merge receive over a float8 key%s", "", 0);
	}
	p.nkeys = nkeys;
	p.n = rel->nrows;
	if (rel->nrows < 2)
		return CBGPU_OK;
	blocks = (int) ((rel->nrows + 255) / 256);
	if (blocks > ctx->sm_count * 8)
		blocks = ctx->sm_count * 8;
	CB_CUDA(ctx, cudaMallocAsync(&p.nfound, sizeof(int32_t) + MERGE_MAXRUNS * sizeof(unsigned long long) + 8, ctx->stream));
	p.found = (unsigned long long *) ((char *) p.nfound + 8);
	CB_CUDA(ctx, cudaMemsetAsync(p.nfound, 0, sizeof(int32_t), ctx->stream));
	k_merge_find_runs<<<blocks, 256, 0, ctx->stream>>>(p);
	CB_LAUNCHED(ctx, "k_merge_find_runs");
	CB_CUDA(ctx, cudaMemcpyAsync(&nfound, p.nfound, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
	CB_CUDA(ctx, cudaMemcpyAsync(found, p.found, sizeof(found), cudaMemcpyDeviceToHost, ctx->stream));
	CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	CB_CUDA(ctx, cudaFreeAsync(p.nfound, ctx->stream));
	if (nfound + 1 > max_runs)
		return cb_fail(ctx, CBGPU_ERR_INVALID, "merge receive: %s%lld sorted runs arrived from fewer senders - a sender's stream is not in the Motion's sort order",
					   "", (long long) nfound + 1);
	if (nruns_out)
		*nruns_out = nfound + 1;
	if (nfound == 0)
		return CBGPU_OK;		/* one run: arrival order is the order */
	/* run starts in ascending order (found in atomic order) */
	for (int a = 1; a < nfound; a++)
		for (int b = a; b > 0 && found[b] < found[b - 1]; b--)
		{
			unsigned long long t = found[b];

			found[b] = found[b - 1];
			found[b - 1] = t;
		}
	p.nruns = nfound + 1;
	p.start[0] = 0;
	for (int a = 0; a < nfound; a++)
		p.start[a + 1] = (int64_t) found[a];
	p.start[p.nruns] = rel->nrows;
	CB_CUDA(ctx, cudaMallocAsync(&p.order, (size_t) rel->nrows * sizeof(uint32_t), ctx->stream));
	k_merge_place<<<blocks, 256, 0, ctx->stream>>>(p);
	CB_LAUNCHED(ctx, "k_merge_place");
	*order_dev = p.order;
	return CBGPU_OK;
}

extern "C" int
cbgpu_topn(cbgpu_ctx *ctx, cbgpu_rel *rel, const int32_t *keycols, const int32_t *descending, const int32_t *unsigned_cmp,
		   int32_t nkeys, int64_t limit, uint32_t *host_idx, int64_t *nout)
{
	TopnParams	p;
	uint32_t   *cand1 = NULL,
			   *cand2 = NULL;
	int64_t		n1,
				n2;

	if (nkeys < 1 || nkeys > TOPN_MAXKEYS || limit < 1 || limit > TOPN_MAXK)
		return cb_fail(ctx, CBGPU_ERR_UNSUPPORTED, "device top-N supports 1..4 sort keys and LIMIT <= 64 (%s got %lld)", "", limit);
	memset(&p, 0, sizeof(p));
	for (int i = 0; i < nkeys; i++)
	{
		if (keycols[i] < 0 || keycols[i] >= rel->ncols)
			return cb_fail(ctx, CBGPU_ERR_INVALID, "cbgpu_topn: bad key column%s %lld", "", keycols[i]);
		if (rel->nulls[keycols[i]])
			return cb_fail(ctx, CBGPU_ERR_UNSUPPORTED, "device top-N over a nullable sort key (%s column %lld)", "", keycols[i]);
		p.key[i] = rel->data[keycols[i]];
		p.keytype[i] = rel->types[keycols[i]];
		p.desc[i] = descending[i];
		p.uns[i] = unsigned_cmp ? unsigned_cmp[i] : 0;
		if (p.keytype[i] == CB_FLOAT8)
			return cb_fail(ctx, CBGPU_ERR_UNSUPPORTED, "device top-N over a float8 key%s", "", 0);
	}
	p.nkeys = nkeys;
	p.k = (int32_t) limit;
	*nout = 0;
	if (rel->nrows == 0)
		return CBGPU_OK;
	{
		/* threshold selection first (see above); the tournament below only when it cannot decide */
		unsigned long long *scratch = NULL;		/* [TOPN_BLOCKS] minima, [TOPN_BLOCKS] maxima, [3] threshold / min / max */
		uint32_t   *cand = NULL;
		int		   *ncand = NULL;
		unsigned long long h3[3];
		int			hn = 0;
		bool		done = false;

		p.in_idx = NULL;
		p.nin = rel->nrows;
		CB_CUDA(ctx, cudaMallocAsync(&cand1, (TOPN_MAXK + 1) * sizeof(uint32_t), ctx->stream));
		if (rel->nrows <= TOPN_CAND)
		{
			k_topn_final<<<1, 1024, 0, ctx->stream>>>(p, NULL, NULL, rel->nrows, cand1);
			CB_LAUNCHED(ctx, "k_topn_final");
			done = true;
		}
		else
		{
			int			nblocks = (int) ((rel->nrows + 63) / 64);	/* > TOPN_CAND rows: at least TOPN_MAXK blocks */

			if (nblocks > TOPN_BLOCKS)
				nblocks = TOPN_BLOCKS;
			CB_CUDA(ctx, cudaMallocAsync(&scratch, (2 * TOPN_BLOCKS + 3) * sizeof(unsigned long long), ctx->stream));
			CB_CUDA(ctx, cudaMallocAsync(&cand, TOPN_CAND * sizeof(uint32_t), ctx->stream));
			CB_CUDA(ctx, cudaMallocAsync(&ncand, sizeof(int), ctx->stream));
			for (int keyi = 0; keyi < nkeys && !done; keyi++)
			{
				k_topn_min<<<nblocks, 256, 0, ctx->stream>>>(p, keyi, scratch, scratch + TOPN_BLOCKS);
				CB_LAUNCHED(ctx, "k_topn_min");
				k_topn_kth<<<1, TOPN_BLOCKS, 0, ctx->stream>>>(scratch, scratch + TOPN_BLOCKS, nblocks, p.k, scratch + 2 * TOPN_BLOCKS);
				CB_LAUNCHED(ctx, "k_topn_kth");
				CB_CUDA(ctx, cudaMemcpyAsync(h3, scratch + 2 * TOPN_BLOCKS, sizeof(h3), cudaMemcpyDeviceToHost, ctx->stream));
				CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
				if (h3[1] == h3[2])
					continue;	/* a constant key (e.g. the high half of small 128-bit sums) decides nothing */
				CB_CUDA(ctx, cudaMemsetAsync(ncand, 0, sizeof(int), ctx->stream));
				k_topn_filter<<<nblocks, 256, 0, ctx->stream>>>(p, keyi, scratch + 2 * TOPN_BLOCKS, cand, ncand);
				CB_LAUNCHED(ctx, "k_topn_filter");
				CB_CUDA(ctx, cudaMemcpyAsync(&hn, ncand, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
				CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
				if (hn <= TOPN_CAND)
				{
					k_topn_final<<<1, 1024, 0, ctx->stream>>>(p, cand, ncand, 0, cand1);
					CB_LAUNCHED(ctx, "k_topn_final");
					done = true;
				}
				break;			/* heavily tied key: the tournament sorts it out */
			}
			CB_CUDA(ctx, cudaFreeAsync(scratch, ctx->stream));
			CB_CUDA(ctx, cudaFreeAsync(cand, ctx->stream));
			CB_CUDA(ctx, cudaFreeAsync(ncand, ctx->stream));
		}
		if (done)
			goto fetch;
		CB_CUDA(ctx, cudaFreeAsync(cand1, ctx->stream));
		cand1 = NULL;
	}
	/* tournament: every pass gives each thread ~64 candidates and keeps its best k, until one thread
	 * holds the answer; passes shrink the candidate list by ~64 / k each */
	{
		const int	threads = 128;
		const int64_t per_thread = 64;
		const uint32_t *in = NULL;
		int64_t		nin = rel->nrows;
		int			which = 0;
		int64_t		cap = ((rel->nrows + per_thread - 1) / per_thread + threads) * limit;

		CB_CUDA(ctx, cudaMallocAsync(&cand1, cap * sizeof(uint32_t), ctx->stream));
		CB_CUDA(ctx, cudaMallocAsync(&cand2, cap * sizeof(uint32_t), ctx->stream));
		for (;;)
		{
			int64_t		nthreads = (nin + per_thread - 1) / per_thread;
			int			blocks = (int) ((nthreads + threads - 1) / threads);
			int			tpb = nthreads < threads ? (int) nthreads : threads;
			uint32_t   *outbuf = which ? cand2 : cand1;

			if (nthreads <= 1)
			{
				blocks = 1;
				tpb = 1;
			}
			p.in_idx = in;
			p.nin = nin;
			p.out_idx = outbuf;
			k_topn<<<blocks, tpb, 0, ctx->stream>>>(p);
			CB_LAUNCHED(ctx, "k_topn");
			n1 = (int64_t) blocks * tpb * limit;
			if (blocks == 1 && tpb == 1)
			{
				if (outbuf != cand1)
					CB_CUDA(ctx, cudaMemcpyAsync(cand1, outbuf, limit * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
				break;
			}
			in = outbuf;
			nin = n1;
			which ^= 1;
		}
		(void) n2;
	}
fetch:
	CB_CUDA(ctx, cudaMemcpyAsync(host_idx, cand1, limit * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
	CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	cudaFreeAsync(cand1, ctx->stream);
	cudaFreeAsync(cand2, ctx->stream);
	int64_t		n = 0;

	while (n < limit && host_idx[n] != 0xFFFFFFFFu)
		n++;
	*nout = n;
	return CBGPU_OK;
}
