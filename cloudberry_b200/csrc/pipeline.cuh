/*
 * pipeline.cuh - device-side form of a pipeline descriptor and the sink primitives shared by the
 * generic and the pattern-specialised pipeline kernels.
 */
#pragma once
#include "common.cuh"

#define CBP_MAX_STAGES 12

/* internal fused opcodes produced by the host-side peephole pass (never part of the C ABI) */
#define XOP_FILTER_COL 100		/* a = col | cmp << 16, imm = constant                               */
#define XOP_PROBE_COLS 101		/* a = probe, imm = key column indexes, 8 bits each                  */
#define XOP_MULCSUB 102			/* a = colA | colB << 16, imm = k: push colA * (k - colB)            */

struct DProbe
{
	HtDev		ht;
	int32_t		jointype;
	int32_t		nkeys;
	int32_t		keytype[CBP_MAX_KEYS];
	const uint32_t *keydict[CBP_MAX_KEYS];
	int32_t		null_key_drops;
};

struct DSink
{
	int32_t		kind;
	AggDev		agg;
	int32_t		nkeys;
	int32_t		keytype[CBP_MAX_KEYS];
	const uint32_t *keydict[CBP_MAX_KEYS];
	int32_t		naccs;
	CbpAcc		accs[CBP_MAX_AGGS];
	int32_t		nout;
	void	   *outcol[CBP_MAX_OUT];
	uint8_t    *outnull[CBP_MAX_OUT];
	int32_t		outtype[CBP_MAX_OUT];
	unsigned long long *out_count;
	int64_t		out_capacity;
	int32_t		nhash;
	int32_t		hashtype[CBP_MAX_KEYS];
	const uint32_t *hashdict[CBP_MAX_KEYS];
	int32_t		nsegs;
	int64_t		seg_capacity;
	void *const *part_cols;		/* direct Motion: destination d's column c at part_cols[d * nout + c]  */
	unsigned long long *const *part_counts;	/* ... and its row counter (peer memory)               */
	uint8_t *const *part_nulls;	/* ... and its NULL bytes                                             */
	unsigned long long part_nullmask;
	int32_t    *part_flags;		/* a full destination ORs CBGPU_DX_OVERFLOW in here (NULL: status word) */
	int64_t		seg_base[64];	/* staged: first row of destination d's range in the output columns    */
	int64_t		seg_cap[64];	/* ... and how many rows fit                                           */
};

struct PipeDev
{
	int64_t		nrows;
	const uint8_t *visimap;
	int32_t		drv_nsrc;
	const uint32_t *drv_idx[CBP_MAX_SRC];
	int32_t		ncols;
	CbpColumn	cols[CBP_MAX_COLS];
	int32_t		nops;
	CbpOp		ops[CBP_MAX_OPS];
	int32_t		nprobes;
	int32_t		src_base;		/* probe j's inner rows are source src_base + j                      */
	int32_t		nsrc;			/* sources in use                                                    */
	int32_t		nstages;		/* program stages (cut after every FILTER / PROBE)                   */
	int32_t		stage_pc[CBP_MAX_STAGES + 2];
	int32_t		q_off[CBP_MAX_STAGES];	/* per-warp queue b: word offset and words per entry               */
	int32_t		q_ew[CBP_MAX_STAGES];
	int32_t		q_words;		/* words of queue space per warp                                     */
	DProbe		probes[CBP_MAX_SRC - 1];
	DSink		sink;
	int		   *status;
};

/* one accumulator update of advance_aggregates (backend/executor/nodeAgg.c:856).  `args` are the
 * stack values after the keys; argnull their NULL bits. */
__device__ __forceinline__ void
sink_acc_update(const AggDev &t, int slot, int a, const CbpAcc &acc, const int64_t *args, uint64_t argnull)
{
	int64_t    *np = t.n + (size_t) slot * t.naccs + a;
	unsigned long long *sp = t.sum + ((size_t) slot * t.naccs + a) * 2;

	switch (acc.kind)
	{
		case CBP_ACC_COUNT:		/* int8inc / int8inc_any (utils/adt/int8.c:805) */
			if (acc.arg < 0 || !((argnull >> acc.arg) & 1))
				atomicAdd((unsigned long long *) np, 1ull);
			break;
		case CBP_ACC_SUM_INT:	/* int4_sum / int8_avg_accum / numeric_avg_accum: N += 1, sumX += v */
			if (!((argnull >> acc.arg) & 1))
			{
				atomicAdd((unsigned long long *) np, 1ull);
				atomic_add128_signed(sp, args[acc.arg]);
			}
			break;
		case CBP_ACC_SUM_FLOAT:	/* float8pl / float8_accum Sx (tree order here: tolerance, not bit-exact) */
			if (!((argnull >> acc.arg) & 1))
			{
				atomicAdd((unsigned long long *) np, 1ull);
				atomicAdd((double *) sp, __longlong_as_double(args[acc.arg]));
			}
			break;
		case CBP_ACC_MIN:
			if (!((argnull >> acc.arg) & 1))
			{
				atomicAdd((unsigned long long *) np, 1ull);
				atomicMin((long long *) sp, (long long) args[acc.arg]);
			}
			break;
		case CBP_ACC_MAX:
			if (!((argnull >> acc.arg) & 1))
			{
				atomicAdd((unsigned long long *) np, 1ull);
				atomicMax((long long *) sp, (long long) args[acc.arg]);
			}
			break;
		case CBP_ACC_MERGE_COUNT:	/* int8pl over partial counts */
			if (!((argnull >> acc.arg) & 1))
				atomicAdd((unsigned long long *) np, (unsigned long long) args[acc.arg]);
			break;
		case CBP_ACC_MERGE_INT:	/* int8_avg_combine / numeric_avg_combine (numeric.c:5726, 4946) */
			if (!((argnull >> acc.arg) & 1))
			{
				atomicAdd((unsigned long long *) np, (unsigned long long) args[acc.arg]);
				atomic_add128(sp, (unsigned long long) args[acc.arg + 1], (unsigned long long) args[acc.arg + 2]);
			}
			break;
		case CBP_ACC_MERGE_MIN:
		case CBP_ACC_MERGE_MAX:
			if (!((argnull >> acc.arg) & 1) && args[acc.arg] > 0)
			{
				atomicAdd((unsigned long long *) np, (unsigned long long) args[acc.arg]);
				if (acc.kind == CBP_ACC_MERGE_MIN)
					atomicMin((long long *) sp, (long long) args[acc.arg + 1]);
				else
					atomicMax((long long *) sp, (long long) args[acc.arg + 1]);
			}
			break;
		case CBP_ACC_MERGE_FLOAT:	/* float8_combine (float.c:2886): N and Sx add */
			if (!((argnull >> acc.arg) & 1))
			{
				atomicAdd((unsigned long long *) np, (unsigned long long) args[acc.arg]);
				atomicAdd((double *) sp, __longlong_as_double(args[acc.arg + 1]));
			}
			break;
	}
}

__device__ __forceinline__ void
sink_store(void *col, int type, uint64_t pos, int64_t v)
{
	switch (type)
	{
		case CB_INT4: case CB_DATE: case CB_DICT32:
			((int32_t *) col)[pos] = (int32_t) v;
			break;
		case CB_INT8: case CB_NUMERIC: case CB_FLOAT8:
			((int64_t *) col)[pos] = v;
			break;
		default:
			((uint8_t *) col)[pos] = (uint8_t) v;
	}
}

void		cb_agg_touch(cbgpu_aggtable *t);
int			cb_klog_begin(cbgpu_ctx *ctx, const char *name);
void		cb_klog_end(cbgpu_ctx *ctx, int i);
int			cb_pipeline_to_dev(cbgpu_ctx *ctx, const CbPipeline *p, PipeDev *d);
int			cb_try_probe_chain(cbgpu_ctx *ctx, const CbPipeline *p, const PipeDev *d, bool *handled);
int			cb_try_specialised(cbgpu_ctx *ctx, const CbPipeline *p, const PipeDev *d, bool *handled);
