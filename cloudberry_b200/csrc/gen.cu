/*
 * gen.cu - synthetic TPC-H shaped tables generated directly in HBM (bench / test harness support).
 *
 * Counter based: every value is a pure function of (seed, column id, row index) - the same
 * formulas as cloudberry_b200/tpch.py (gen_lineitem, gen_orders, ...), which the CPU oracle uses,
 * so host and device tables are identical row for row and any row range can be regenerated
 * independently (SURVEY.md 8d "synthetic inputs": dbgen-like value domains, sparse order keys,
 * 1-7 lines per order, flags derived from dates).
 */
#include "common.cuh"

__host__ __device__ __forceinline__ uint64_t
gen_mix(uint64_t x)
{
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return x ^ (x >> 31);
}

__host__ __device__ __forceinline__ uint64_t
gen_u(uint64_t seed, uint64_t col, uint64_t idx)
{
	return gen_mix(seed * 0x9E3779B97F4A7C15ull + col * 0xD1B54A32D192ED03ull + idx);
}

#define GEN_STARTDATE (-2922)	/* 1992-01-01 in days since 2000-01-01 */
#define GEN_CURRENTDATE (-1659) /* 1995-06-17 */
#define GEN_DATE_SPAN 2406

__host__ __device__ __forceinline__ int64_t
gen_order_key(int64_t idx)
{
	return (idx / 8) * 32 + (idx % 8) + 1;
}

struct GenLineitem
{
	int64_t    *orderkey;
	int32_t    *suppkey;
	int64_t    *quantity, *extendedprice, *discount, *tax;
	uint8_t    *returnflag, *linestatus;
	int32_t    *shipdate;
	int64_t		n;
	int64_t		row_lo;
	uint64_t	seed;
	uint64_t	n_supp, n_part;
};

__global__ void
k_gen_lineitem(GenLineitem g)
{
	/* 28 lines per block of 7 orders: order k of the block has k + 1 lines */
	const int	off28[28] = {0, 1, 1, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5, 5, 6, 6, 6, 6, 6, 6, 6};
	int64_t		i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	int64_t		stride = (int64_t) gridDim.x * blockDim.x;

	for (; i < g.n; i += stride)
	{
		int64_t		j = g.row_lo + i;
		int64_t		oidx = (j / 28) * 7 + off28[j % 28];
		int64_t		odate = GEN_STARTDATE + (int64_t) (gen_u(g.seed, 12, (uint64_t) oidx) % GEN_DATE_SPAN);
		int64_t		qty = 1 + (int64_t) (gen_u(g.seed, 21, (uint64_t) j) % 50);
		int64_t		pk = 1 + (int64_t) (gen_u(g.seed, 22, (uint64_t) j) % g.n_part);
		int64_t		price = 90000 + (pk / 10) % 20001 + 100 * (pk % 1000);
		int64_t		ship = odate + 1 + (int64_t) (gen_u(g.seed, 25, (uint64_t) j) % 121);
		int64_t		receipt = ship + 1 + (int64_t) (gen_u(g.seed, 26, (uint64_t) j) % 30);
		uint8_t		ra = (gen_u(g.seed, 27, (uint64_t) j) % 2 == 0) ? 'R' : 'A';

		g.orderkey[i] = gen_order_key(oidx);
		g.suppkey[i] = (int32_t) (1 + gen_u(g.seed, 23, (uint64_t) j) % g.n_supp);
		g.quantity[i] = qty * 100;
		g.extendedprice[i] = qty * price;
		g.discount[i] = (int64_t) (gen_u(g.seed, 24, (uint64_t) j) % 11);
		g.tax[i] = (int64_t) (gen_u(g.seed, 28, (uint64_t) j) % 9);
		g.returnflag[i] = receipt <= GEN_CURRENTDATE ? ra : (uint8_t) 'N';
		g.linestatus[i] = ship > GEN_CURRENTDATE ? 'O' : 'F';
		g.shipdate[i] = (int32_t) ship;
	}
}

static int
gen_check(cbgpu_ctx *ctx, cbgpu_rel *rel, const int *types, int ncols, const char *name)
{
	if (rel->ncols != ncols)
		return cb_fail(ctx, CBGPU_ERR_INVALID, "generator: relation %s has the wrong column count (%lld)", name, rel->ncols);
	for (int i = 0; i < ncols; i++)
		if (rel->types[i] != types[i])
			return cb_fail(ctx, CBGPU_ERR_INVALID, "generator: relation %s column %lld has the wrong type", name, i);
	return CBGPU_OK;
}

static int
gen_blocks(cbgpu_ctx *ctx, int64_t n)
{
	int64_t		b = (n + 255) / 256;

	if (b > (int64_t) ctx->sm_count * 16)
		b = (int64_t) ctx->sm_count * 16;
	return b < 1 ? 1 : (int) b;
}

extern "C" int
cbgpu_gen_lineitem(cbgpu_ctx *ctx, cbgpu_rel *rel, uint64_t seed, int64_t row_lo, int64_t n_supp, int64_t n_part)
{
	const int	types[9] = {CB_INT8, CB_INT4, CB_NUMERIC, CB_NUMERIC, CB_NUMERIC, CB_NUMERIC, CB_BPCHAR1, CB_BPCHAR1, CB_DATE};
	int			rc = gen_check(ctx, rel, types, 9, "lineitem");
	GenLineitem g;

	if (rc)
		return rc;
	g.orderkey = (int64_t *) rel->data[0];
	g.suppkey = (int32_t *) rel->data[1];
	g.quantity = (int64_t *) rel->data[2];
	g.extendedprice = (int64_t *) rel->data[3];
	g.discount = (int64_t *) rel->data[4];
	g.tax = (int64_t *) rel->data[5];
	g.returnflag = (uint8_t *) rel->data[6];
	g.linestatus = (uint8_t *) rel->data[7];
	g.shipdate = (int32_t *) rel->data[8];
	g.n = rel->nrows;
	g.row_lo = row_lo;
	g.seed = seed;
	g.n_supp = (uint64_t) n_supp;
	g.n_part = (uint64_t) n_part;
	if (rel->nrows == 0)
		return CBGPU_OK;
	k_gen_lineitem<<<gen_blocks(ctx, rel->nrows), 256, 0, ctx->stream>>>(g);
	CB_LAUNCHED(ctx, "k_gen_lineitem");
	return CBGPU_OK;
}

__global__ void
k_gen_orders(int64_t *orderkey, int32_t *custkey, int32_t *orderdate, int32_t *shipprio, int64_t n, int64_t row_lo,
			 uint64_t seed, uint64_t n_cust)
{
	int64_t		i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	int64_t		stride = (int64_t) gridDim.x * blockDim.x;

	for (; i < n; i += stride)
	{
		int64_t		o = row_lo + i;

		orderkey[i] = gen_order_key(o);
		custkey[i] = (int32_t) (1 + gen_u(seed, 11, (uint64_t) o) % n_cust);
		orderdate[i] = (int32_t) (GEN_STARTDATE + (int64_t) (gen_u(seed, 12, (uint64_t) o) % GEN_DATE_SPAN));
		shipprio[i] = 0;
	}
}

extern "C" int
cbgpu_gen_orders(cbgpu_ctx *ctx, cbgpu_rel *rel, uint64_t seed, int64_t row_lo, int64_t n_cust)
{
	const int	types[4] = {CB_INT8, CB_INT4, CB_DATE, CB_INT4};
	int			rc = gen_check(ctx, rel, types, 4, "orders");

	if (rc)
		return rc;
	if (rel->nrows == 0)
		return CBGPU_OK;
	k_gen_orders<<<gen_blocks(ctx, rel->nrows), 256, 0, ctx->stream>>>((int64_t *) rel->data[0], (int32_t *) rel->data[1],
																		  (int32_t *) rel->data[2], (int32_t *) rel->data[3],
																		  rel->nrows, row_lo, seed, (uint64_t) n_cust);
	CB_LAUNCHED(ctx, "k_gen_orders");
	return CBGPU_OK;
}

__global__ void
k_gen_keyed(int32_t *key, int32_t *nation, uint8_t *seg, int64_t n, int64_t row_lo, uint64_t seed, uint64_t col_nation,
			uint64_t col_seg)
{
	int64_t		i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	int64_t		stride = (int64_t) gridDim.x * blockDim.x;

	for (; i < n; i += stride)
	{
		const int64_t r = row_lo + i;

		key[i] = (int32_t) (r + 1);
		nation[i] = (int32_t) (gen_u(seed, col_nation, (uint64_t) r) % 25);
		if (seg)
			seg[i] = (uint8_t) (gen_u(seed, col_seg, (uint64_t) r) % 5);
	}
}

extern "C" int
cbgpu_gen_customer(cbgpu_ctx *ctx, cbgpu_rel *rel, uint64_t seed)
{
	return cbgpu_gen_customer_range(ctx, rel, seed, 0);
}

extern "C" int
cbgpu_gen_supplier(cbgpu_ctx *ctx, cbgpu_rel *rel, uint64_t seed)
{
	return cbgpu_gen_supplier_range(ctx, rel, seed, 0);
}

extern "C" int
cbgpu_gen_customer_range(cbgpu_ctx *ctx, cbgpu_rel *rel, uint64_t seed, int64_t row_lo)
{
	const int	types[3] = {CB_INT4, CB_INT4, CB_DICT8};
	int			rc = gen_check(ctx, rel, types, 3, "customer");

	if (rc)
		return rc;
	if (rel->nrows == 0)
		return CBGPU_OK;
	k_gen_keyed<<<gen_blocks(ctx, rel->nrows), 256, 0, ctx->stream>>>((int32_t *) rel->data[0], (int32_t *) rel->data[1],
																		 (uint8_t *) rel->data[2], rel->nrows, row_lo, seed, 31, 32);
	CB_LAUNCHED(ctx, "k_gen_keyed");
	return CBGPU_OK;
}

extern "C" int
cbgpu_gen_supplier_range(cbgpu_ctx *ctx, cbgpu_rel *rel, uint64_t seed, int64_t row_lo)
{
	const int	types[2] = {CB_INT4, CB_INT4};
	int			rc = gen_check(ctx, rel, types, 2, "supplier");

	if (rc)
		return rc;
	if (rel->nrows == 0)
		return CBGPU_OK;
	k_gen_keyed<<<gen_blocks(ctx, rel->nrows), 256, 0, ctx->stream>>>((int32_t *) rel->data[0], (int32_t *) rel->data[1],
																		 NULL, rel->nrows, row_lo, seed, 41, 0);
	CB_LAUNCHED(ctx, "k_gen_keyed");
	return CBGPU_OK;
}

/* ---------------------------------------------------------------------------------------------
 * Star Schema Benchmark fact table (BASELINE.json configs[4]): the formulas of cloudberry_b200/ssb.py
 * gen_tables, rows [row_lo, row_lo + n) of lineorder.  Dimensions are small and come from the host.
 * --------------------------------------------------------------------------------------------- */
struct GenLineorder
{
	int32_t    *custkey, *partkey, *suppkey, *orderdate;
	int64_t    *revenue, *supplycost;
	int64_t		n;
	int64_t		row_lo;
	uint64_t	seed;
	uint64_t	n_cust, n_part, n_supp;
};

/* d_datekey (yyyymmdd) of day `idx` counted from 1992-01-01; 1992 and 1996 are the leap years in range */
__host__ __device__ __forceinline__ int32_t
gen_ssb_datekey(int idx)
{
	const int	mdays[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
	int			y = 1992;

	for (;;)
	{
		const int	ylen = (y % 4 == 0) ? 366 : 365;

		if (idx < ylen)
			break;
		idx -= ylen;
		y++;
	}
	for (int m = 0; m < 12; m++)
	{
		const int	n = mdays[m] + ((m == 1 && y % 4 == 0) ? 1 : 0);

		if (idx < n)
			return y * 10000 + (m + 1) * 100 + idx + 1;
		idx -= n;
	}
	return y * 10000 + 1231;
}

__global__ void
k_gen_lineorder(GenLineorder g)
{
	int64_t		i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
	int64_t		stride = (int64_t) gridDim.x * blockDim.x;

	for (; i < g.n; i += stride)
	{
		const uint64_t j = (uint64_t) (g.row_lo + i);

		g.custkey[i] = (int32_t) (1 + gen_u(g.seed, 51, j) % g.n_cust);
		g.partkey[i] = (int32_t) (1 + gen_u(g.seed, 52, j) % g.n_part);
		g.suppkey[i] = (int32_t) (1 + gen_u(g.seed, 53, j) % g.n_supp);
		g.orderdate[i] = gen_ssb_datekey((int) (gen_u(g.seed, 54, j) % 2556));
		g.revenue[i] = (int64_t) (100 + gen_u(g.seed, 55, j) % 10000000);
		g.supplycost[i] = (int64_t) (50 + gen_u(g.seed, 56, j) % 120000);
	}
}

extern "C" int
cbgpu_gen_ssb_lineorder(cbgpu_ctx *ctx, cbgpu_rel *rel, uint64_t seed, int64_t row_lo, int64_t n_cust, int64_t n_part, int64_t n_supp)
{
	const int	types[6] = {CB_INT4, CB_INT4, CB_INT4, CB_INT4, CB_INT8, CB_INT8};
	int			rc = gen_check(ctx, rel, types, 6, "lineorder");
	GenLineorder g;

	if (rc)
		return rc;
	if (n_cust < 1 || n_part < 1 || n_supp < 1)
		return cb_fail(ctx, CBGPU_ERR_INVALID, "generator: lineorder needs the dimension sizes%s", "", 0);
	g.custkey = (int32_t *) rel->data[0];
	g.partkey = (int32_t *) rel->data[1];
	g.suppkey = (int32_t *) rel->data[2];
	g.orderdate = (int32_t *) rel->data[3];
	g.revenue = (int64_t *) rel->data[4];
	g.supplycost = (int64_t *) rel->data[5];
	g.n = rel->nrows;
	g.row_lo = row_lo;
	g.seed = seed;
	g.n_cust = (uint64_t) n_cust;
	g.n_part = (uint64_t) n_part;
	g.n_supp = (uint64_t) n_supp;
	CB_CUDA(ctx, cudaSetDevice(ctx->device));
	if (rel->nrows > 0)
	{
		k_gen_lineorder<<<gen_blocks(ctx, rel->nrows), 256, 0, ctx->stream>>>(g);
		CB_LAUNCHED(ctx, "k_gen_lineorder");
	}
	return CBGPU_OK;
}
