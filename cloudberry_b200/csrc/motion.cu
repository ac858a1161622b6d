/*
 * motion.cu - the interconnect: Redistribute / Gather / Broadcast Motion between GPU-segments over
 * NCCL (NVLink 5 / NVSwitch), one process per GPU.
 *
 * Replaces the reference's tuple-at-a-time path: doSendTuple -> SendTuple -> SerializeTuple ->
 * MotionIPCLayer->SendTupleChunkToAMS on the sender (backend/executor/nodeMotion.c:1181,
 * backend/cdb/motion/cdbmotion.c:425, tupser.c:349) and RecvTupleFrom -> processIncomingChunks ->
 * CvtChunksToTup on the receiver (cdbmotion.c:549,628; tupser.c:519), with their <= 8 KB chunks,
 * acks and flow control (contrib/interconnect/udp/ic_udpifc.c).  Here the sender slice's kernel has
 * already hashed (cdbhash + jump consistent hash) and scattered its rows into per-destination
 * contiguous column ranges (CBP_SINK_PARTITION); what is left is a count exchange (one small
 * all-gather, so skewed destinations are sized exactly) and one grouped ncclSend/ncclRecv per
 * column - an all-to-all whose every byte is payload.  End of stream is implicit in the counts
 * (the reference sends TC_END_OF_STREAM chunks, tupchunk.h:23-28).
 */
#include "common.cuh"

#include <nccl.h>
#include <stdlib.h>
#include <string.h>

struct cbgpu_motion
{
	cbgpu_ctx  *ctx;
	ncclComm_t	comm;
	int			rank;
	int			nranks;
	long long  *d_counts;		/* [nranks * nranks] scratch for the count exchange                   */
	int64_t		bytes_sent;		/* payload bytes this rank handed to NCCL (diagnostics / bench)       */
	int64_t		exchanges;
	/* peer-memory window: one cudaMalloc'ed arena per rank, mapped into every other rank's process
	 * (CUDA IPC), so a sender slice's PARTITION sink stores rows straight into the receiver's HBM
	 * over NVLink - no staging buffer, no payload through NCCL (direct Redistribute, below) */
	char	   *win;
	size_t		win_bytes;		/* the smallest window of all ranks                                   */
	char	   *peer_win[64];	/* peer_win[rank] == win                                              */
	bool		direct_ok;
	void	  **d_tab;			/* device table handed to the sink: [64 * CBP_MAX_OUT] column bases, [64] counters */
	void	  **h_tab;			/* pinned host copy                                                   */
	/* the direct exchange in flight */
	int32_t		dx_ncols;
	int32_t		dx_types[CBP_MAX_OUT];
	int32_t		dx_dscales[CBP_MAX_OUT];
	int64_t		dx_cap;
	size_t		dx_off[CBP_MAX_OUT];
	int64_t		direct_bytes;	/* payload bytes stored into peers' windows (diagnostics / bench)      */
	int64_t		direct_exchanges;
	int64_t		gather_seq;		/* direct Gathers so far: picks the buffer                            */
};

#define DX_TAB_COLS (64 * CBP_MAX_OUT)
#define DX_ALIGN 256
/* the tail of every window: two alternating Gather buffers; each holds [64 row counts] + one payload
 * slot per sender */
#define GX_BYTES ((size_t) 32 << 20)
#define GX_HDR 1024

#define CB_NCCL(ctx, call) \
	do { \
		ncclResult_t r__ = (call); \
		if (r__ != ncclSuccess) \
		{ \
			snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: %s: %s", __FILE__, __LINE__, #call, ncclGetErrorString(r__)); \
			return CBGPU_ERR_CUDA; \
		} \
	} while (0)

extern "C" int
cbgpu_motion_unique_id(void *out128)
{
	ncclUniqueId id;

	if (sizeof(id) != 128)
		return CBGPU_ERR_INVALID;
	if (ncclGetUniqueId(&id) != ncclSuccess)
		return CBGPU_ERR_CUDA;
	memcpy(out128, &id, 128);
	return CBGPU_OK;
}

static int	motion_window_setup(cbgpu_motion *m);
static void motion_window_teardown(cbgpu_motion *m);

extern "C" int
cbgpu_motion_create(cbgpu_ctx *ctx, int rank, int nranks, const void *unique_id128, cbgpu_motion **out)
{
	cbgpu_motion *m = (cbgpu_motion *) calloc(1, sizeof(cbgpu_motion));
	ncclUniqueId id;

	if (!m)
		return CBGPU_ERR_NOMEM;
	memcpy(&id, unique_id128, 128);
	m->ctx = ctx;
	m->rank = rank;
	m->nranks = nranks;
	CB_CUDA(ctx, cudaSetDevice(ctx->device));
	CB_NCCL(ctx, ncclCommInitRank(&m->comm, nranks, id, rank));
	CB_CUDA(ctx, cudaMalloc(&m->d_counts, sizeof(long long) * (size_t) (nranks + 2) * (size_t) (nranks + 2)));
	*out = m;
	return motion_window_setup(m);
}

/* ---------------------------------------------------------------------------------------------
 * peer-memory window
 * --------------------------------------------------------------------------------------------- */
struct WinHello
{
	cudaIpcMemHandle_t handle;
	unsigned long long bytes;
	int			ok;
	char		pad[128 - sizeof(cudaIpcMemHandle_t) - sizeof(unsigned long long) - sizeof(int)];
};

/* allocate this rank's window, swap IPC handles through the communicator, map every peer's.
 * Any rank failing at any step turns the direct path off on ALL ranks (the decision is all-reduced):
 * Redistribute then takes the staged NCCL path - both are device paths. */
static int
motion_window_setup(cbgpu_motion *m)
{
	cbgpu_ctx  *ctx = m->ctx;
	const int	n = m->nranks;
	WinHello   *h_all = (WinHello *) calloc((size_t) n, sizeof(WinHello));
	WinHello   *d_all = NULL;
	WinHello	mine;
	size_t		want = (size_t) 16384 << 20;	/* of 180 GB: room for a Motion of ~500 M narrow rows per receiver */
	const char *env = getenv("CBGPU_MOTION_WINDOW_MB");
	int			ok = 1;
	int		   *d_ok = NULL;
	int			h_ok = 0;

	m->direct_ok = false;
	if (!h_all)
		return CBGPU_ERR_NOMEM;
	if (n > 64 || (getenv("CBGPU_MOTION") && strcmp(getenv("CBGPU_MOTION"), "nccl") == 0))
		ok = 0;
	if (env && atoll(env) > 0)
		want = (size_t) atoll(env) << 20;
	memset(&mine, 0, sizeof(mine));
	if (ok)
	{
		size_t		freeb = 0, total = 0;

		cudaMemGetInfo(&freeb, &total);
		while (want > freeb / 2 && want > ((size_t) 64 << 20))
			want >>= 1;
		if (cudaMalloc(&m->win, want) != cudaSuccess || cudaIpcGetMemHandle(&mine.handle, m->win) != cudaSuccess)
		{
			cudaGetLastError();
			ok = 0;
		}
	}
	mine.bytes = want;
	mine.ok = ok;
	CB_CUDA(ctx, cudaMalloc(&d_all, sizeof(WinHello) * (size_t) (n + 1)));
	CB_CUDA(ctx, cudaMemcpyAsync(d_all + n, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
	CB_NCCL(ctx, ncclAllGather(d_all + n, d_all, sizeof(WinHello), ncclInt8, m->comm, ctx->stream));
	CB_CUDA(ctx, cudaMemcpyAsync(h_all, d_all, sizeof(WinHello) * (size_t) n, cudaMemcpyDeviceToHost, ctx->stream));
	CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	m->win_bytes = want;
	for (int p = 0; p < n; p++)
	{
		if (!h_all[p].ok)
			ok = 0;
		if (h_all[p].bytes < m->win_bytes)
			m->win_bytes = (size_t) h_all[p].bytes;
	}
	for (int p = 0; p < n && ok; p++)
	{
		if (p == m->rank)
		{
			m->peer_win[p] = m->win;
			continue;
		}
		if (cudaIpcOpenMemHandle((void **) &m->peer_win[p], h_all[p].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
		{
			cudaGetLastError();
			m->peer_win[p] = NULL;
			ok = 0;
		}
	}
	/* everyone or no one */
	CB_CUDA(ctx, cudaMalloc(&d_ok, sizeof(int)));
	CB_CUDA(ctx, cudaMemcpyAsync(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
	CB_NCCL(ctx, ncclAllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, m->comm, ctx->stream));
	CB_CUDA(ctx, cudaMemcpyAsync(&h_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	cudaFree(d_ok);
	cudaFree(d_all);
	free(h_all);
	if (h_ok)
	{
		CB_CUDA(ctx, cudaMalloc(&m->d_tab, sizeof(void *) * (DX_TAB_COLS + 64)));
		CB_CUDA(ctx, cudaMallocHost(&m->h_tab, sizeof(void *) * (DX_TAB_COLS + 64)));
		m->direct_ok = m->win_bytes >= 4 * GX_BYTES;
	}
	if (!m->direct_ok)
		motion_window_teardown(m);
	return CBGPU_OK;
}

static void
motion_window_teardown(cbgpu_motion *m)
{
	for (int p = 0; p < m->nranks && p < 64; p++)
		if (p != m->rank && m->peer_win[p])
		{
			cudaIpcCloseMemHandle(m->peer_win[p]);
			m->peer_win[p] = NULL;
		}
	if (m->win)
		cudaFree(m->win);
	m->win = NULL;
	if (m->d_tab)
		cudaFree(m->d_tab);
	if (m->h_tab)
		cudaFreeHost(m->h_tab);
	m->d_tab = m->h_tab = NULL;
	m->direct_ok = false;
}

extern "C" void
cbgpu_motion_destroy(cbgpu_motion *m)
{
	if (!m)
		return;
	cudaSetDevice(m->ctx->device);
	cudaStreamSynchronize(m->ctx->stream);
	{
		/* nobody unmaps a window a peer may still be storing into */
		int		   *d = NULL;

		if (m->direct_ok && cudaMalloc(&d, sizeof(int)) == cudaSuccess)
		{
			cudaMemsetAsync(d, 0, sizeof(int), m->ctx->stream);
			ncclAllReduce(d, d, 1, ncclInt32, ncclMin, m->comm, m->ctx->stream);
			cudaStreamSynchronize(m->ctx->stream);
			cudaFree(d);
		}
		motion_window_teardown(m);
	}
	ncclCommDestroy(m->comm);
	cudaFree(m->d_counts);
	free(m);
}

extern "C" int
cbgpu_motion_rank(const cbgpu_motion *m)
{
	return m->rank;
}

extern "C" int
cbgpu_motion_nranks(const cbgpu_motion *m)
{
	return m->nranks;
}

extern "C" int64_t
cbgpu_motion_bytes_sent(const cbgpu_motion *m)
{
	return m->bytes_sent;
}

/* every rank learns every rank's per-destination row counts: matrix[s * nranks + d].  The same
 * all-gather carries which columns of `send` have a NULL map on each rank: a map may exist on one
 * segment only (it follows the data), but sender and receiver must post the same transfers, so every
 * rank gives `send` (zeroed) maps for the union before the rows move. */
static int
exchange_counts(cbgpu_motion *m, cbgpu_rel *send, const int64_t *mine, int64_t *matrix)
{
	cbgpu_ctx  *ctx = m->ctx;
	const int	n = m->nranks;
	long long  *d_all = m->d_counts;
	long long  *d_mine = m->d_counts + (size_t) (n + 1) * n;
	long long	h_mine[65], h_all[65 * 64];
	unsigned long long mask = 0,
				all = 0;

	for (int c = 0; c < send->ncols && c < 64; c++)
		if (send->nulls[c])
			mask |= 1ull << c;
	for (int d = 0; d < n; d++)
		h_mine[d] = mine[d];
	h_mine[n] = (long long) mask;
	CB_CUDA(ctx, cudaMemcpyAsync(d_mine, h_mine, sizeof(long long) * (size_t) (n + 1), cudaMemcpyHostToDevice, ctx->stream));
	CB_NCCL(ctx, ncclAllGather(d_mine, d_all, (size_t) (n + 1), ncclInt64, m->comm, ctx->stream));
	CB_CUDA(ctx, cudaMemcpyAsync(h_all, d_all, sizeof(long long) * (size_t) (n + 1) * n, cudaMemcpyDeviceToHost, ctx->stream));
	if (ctx->trace_on)
		cb_trace_mark(ctx, "nccl:counts");
	CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	for (int s = 0; s < n; s++)
	{
		for (int d = 0; d < n; d++)
			matrix[(size_t) s * n + d] = h_all[(size_t) s * (n + 1) + d];
		all |= (unsigned long long) h_all[(size_t) s * (n + 1) + n];
	}
	for (int c = 0; c < send->ncols && c < 64; c++)
		if (((all >> c) & 1) && !send->nulls[c])
		{
			int			rc = cbgpu_rel_add_nullmap(send, c);

			if (rc)
				return rc;
		}
	return CBGPU_OK;
}

static int
make_recv(cbgpu_motion *m, cbgpu_rel *send, int64_t total, cbgpu_rel **recv)
{
	cbgpu_ctx  *ctx = m->ctx;
	int			rc = cbgpu_rel_create(ctx, total, send->ncols, send->types, send->dscales, recv);

	if (rc)
		return rc;
	for (int c = 0; c < send->ncols; c++)
	{
		if (send->nulls[c])
		{
			rc = cbgpu_rel_add_nullmap(*recv, c);
			if (rc)
				return rc;
		}
		if (send->dict_hash[c])
			cbgpu_rel_share_dict_hash(*recv, c, send, c);
	}
	return CBGPU_OK;
}

/*
 * generic exchange: rank s sends cnt[s][d] rows starting at row off_send[d] of its `send` relation
 * to rank d, which stores them at row off_recv[s] of `recv`.  One grouped send/recv set per column.
 */
static int
exchange_rows(cbgpu_motion *m, cbgpu_rel *send, cbgpu_rel *recv, const int64_t *send_off, const int64_t *send_cnt,
			  const int64_t *recv_off, const int64_t *recv_cnt)
{
	cbgpu_ctx  *ctx = m->ctx;
	int			n = m->nranks;

	/* one NCCL group for the whole exchange: every column's sends and receives fuse into a single
	 * launch */
	CB_NCCL(ctx, ncclGroupStart());
	for (int c = 0; c < send->ncols; c++)
	{
		size_t		w = (size_t) cb_type_w(send->types[c]);

		for (int pass = 0; pass < (send->nulls[c] ? 2 : 1); pass++)
		{
			char	   *sp = pass ? (char *) send->nulls[c] : (char *) send->data[c];
			char	   *rp = pass ? (char *) recv->nulls[c] : (char *) recv->data[c];
			size_t		ww = pass ? 1 : w;

			for (int peer = 0; peer < n; peer++)
			{
				if (peer == m->rank)
					continue;
				if (send_cnt[peer] > 0)
				{
					CB_NCCL(ctx, ncclSend(sp + (size_t) send_off[peer] * ww, (size_t) send_cnt[peer] * ww, ncclInt8, peer, m->comm, ctx->stream));
					m->bytes_sent += send_cnt[peer] * (int64_t) ww;
				}
				if (recv_cnt[peer] > 0)
					CB_NCCL(ctx, ncclRecv(rp + (size_t) recv_off[peer] * ww, (size_t) recv_cnt[peer] * ww, ncclInt8, peer, m->comm, ctx->stream));
			}
		}
	}
	CB_NCCL(ctx, ncclGroupEnd());
	/* rows that stay on this segment never leave the device */
	if (send_cnt[m->rank] > 0)
		for (int c = 0; c < send->ncols; c++)
		{
			size_t		w = (size_t) cb_type_w(send->types[c]);

			CB_CUDA(ctx, cudaMemcpyAsync((char *) recv->data[c] + (size_t) recv_off[m->rank] * w,
										 (char *) send->data[c] + (size_t) send_off[m->rank] * w,
										 (size_t) send_cnt[m->rank] * w, cudaMemcpyDeviceToDevice, ctx->stream));
			if (send->nulls[c])
				CB_CUDA(ctx, cudaMemcpyAsync(recv->nulls[c] + recv_off[m->rank], send->nulls[c] + send_off[m->rank],
											 (size_t) send_cnt[m->rank], cudaMemcpyDeviceToDevice, ctx->stream));
		}
	if (ctx->trace_on)
		cb_trace_mark(ctx, "nccl:rows");
	m->exchanges++;
	return CBGPU_OK;
}

extern "C" int
cbgpu_motion_redistribute(cbgpu_motion *m, cbgpu_rel *send, const int64_t *counts, int64_t seg_capacity, cbgpu_rel **recv)
{
	int			n = m->nranks;
	int64_t    *matrix = (int64_t *) calloc((size_t) n * n, sizeof(int64_t));
	int64_t		send_off[64], send_cnt[64], recv_off[64], recv_cnt[64];
	int64_t		total = 0;
	int			rc;

	if (n > 64)
		return cb_fail(m->ctx, CBGPU_ERR_UNSUPPORTED, "more than 64 segments%s", "", 0);
	rc = exchange_counts(m, send, counts, matrix);
	if (rc == CBGPU_OK)
	{
		for (int s = 0; s < n; s++)
		{
			send_off[s] = (int64_t) s * seg_capacity;
			send_cnt[s] = counts[s];
			recv_off[s] = total;
			recv_cnt[s] = matrix[(size_t) s * n + m->rank];
			total += recv_cnt[s];
		}
		rc = make_recv(m, send, total, recv);
	}
	if (rc == CBGPU_OK)
		rc = exchange_rows(m, send, *recv, send_off, send_cnt, recv_off, recv_cnt);
	free(matrix);
	return rc;
}


/* ---------------------------------------------------------------------------------------------
 * direct Redistribute: partition + exchange as ONE kernel over peer memory.
 *
 *   begin   (collective) every rank announces how many rows enter its sender slice; all derive the
 *           same per-receiver capacity and the same layout of the exchange inside every window
 *           ([row counter][column 0][column 1]...); each rank zeroes ITS counter BEFORE the
 *           announcement all-gather, so no peer can store before the counter is clean; the caller
 *           gets device tables of every destination's column bases and counter
 *   kernel  the sender slice's pipeline runs with its PARTITION sink in direct mode: destination =
 *           cdbhashreduce(keys); a CTA reserves a slice of the DESTINATION's buffer with one
 *           system-scope atomic per destination and run, and stores the rows there over NVLink
 *   end     (collective) one 4-byte all-reduce = "every sender's kernel has finished"; the receiver
 *           reads its counter and moves the rows out of the window into a relation of its own, so
 *           the window is free again when the next Motion's announcement completes
 * --------------------------------------------------------------------------------------------- */
struct CopyCols
{
	void	   *dst[CBP_MAX_OUT];
	const void *src[CBP_MAX_OUT];
	size_t		bytes[CBP_MAX_OUT];
};

__global__ void
k_copy_cols(CopyCols cc)
{
	const int	c = blockIdx.x;
	const size_t n = cc.bytes[c];
	unsigned char *d = (unsigned char *) cc.dst[c];
	const unsigned char *s = (const unsigned char *) cc.src[c];

	for (size_t i = threadIdx.x; i < n; i += blockDim.x)
		d[i] = s[i];
}

extern "C" int
cbgpu_motion_direct_available(const cbgpu_motion *m)
{
	return m->direct_ok ? 1 : 0;
}

extern "C" int64_t
cbgpu_motion_direct_bytes(const cbgpu_motion *m)
{
	return m->direct_bytes;
}

extern "C" int
cbgpu_motion_direct_begin(cbgpu_motion *m, int32_t ncols, const int32_t *types, const int32_t *dscales, int64_t input_rows,
						  int64_t *capacity, void *const **dest_cols, unsigned long long *const **dest_counts)
{
	cbgpu_ctx  *ctx = m->ctx;
	const int	n = m->nranks;
	long long  *d_mine = m->d_counts + (size_t) n * n;
	long long	h_rows[64];
	int64_t		cap = 0,
				total = 0;
	size_t		off = DX_ALIGN;

	if (!m->direct_ok)
		return cb_fail(ctx, CBGPU_ERR_UNSUPPORTED, "direct Redistribute is not available%s (no peer-memory window)", "", 0);
	if (ncols < 1 || ncols > CBP_MAX_OUT)
		return cb_fail(ctx, CBGPU_ERR_INVALID, "direct Redistribute of %s%lld columns", "", ncols);
	/* my counter is clean before anyone can learn that the exchange has started */
	CB_CUDA(ctx, cudaMemsetAsync(m->win, 0, DX_ALIGN, ctx->stream));
	CB_CUDA(ctx, cudaMemcpyAsync(d_mine, &input_rows, sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
	CB_NCCL(ctx, ncclAllGather(d_mine, m->d_counts, 1, ncclInt64, m->comm, ctx->stream));
	CB_CUDA(ctx, cudaMemcpyAsync(h_rows, m->d_counts, sizeof(long long) * (size_t) n, cudaMemcpyDeviceToHost, ctx->stream));
	if (ctx->trace_on)
		cb_trace_mark(ctx, "p2p:announce");
	CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	/* a rank that cannot take part (input_rows < 0: e.g. a nullable column there) vetoes for everyone */
	for (int s = 0; s < n; s++)
		if (h_rows[s] < 0)
			return cb_fail(ctx, CBGPU_ERR_UNSUPPORTED, "direct Redistribute vetoed by segment %s%lld", "", s);
	/* the same arithmetic on every rank: even share + 25 % skew allowance + slack per sender */
	for (int s = 0; s < n; s++)
	{
		total += h_rows[s];
		cap += h_rows[s] / n + h_rows[s] / (4 * n) + 65536;
	}
	if (cap > total)
		cap = total;
	if (cap < 1)
		cap = 1;
	for (int c = 0; c < ncols; c++)
	{
		m->dx_types[c] = types[c];
		m->dx_dscales[c] = dscales ? dscales[c] : 0;
		m->dx_off[c] = off;
		off += (((size_t) cap * (size_t) cb_type_w(types[c])) + DX_ALIGN - 1) / DX_ALIGN * DX_ALIGN;
	}
	if (off + 2 * GX_BYTES > m->win_bytes)
		return cb_fail(ctx, CBGPU_ERR_NOMEM, "direct Redistribute needs %s%lld bytes of window (raise CBGPU_MOTION_WINDOW_MB)", "", (long long) off);
	m->dx_ncols = ncols;
	m->dx_cap = cap;
	for (int d = 0; d < n; d++)
	{
		for (int c = 0; c < ncols; c++)
			m->h_tab[(size_t) d * ncols + c] = m->peer_win[d] + m->dx_off[c];
		m->h_tab[DX_TAB_COLS + d] = m->peer_win[d];
	}
	CB_CUDA(ctx, cudaMemcpyAsync(m->d_tab, m->h_tab, sizeof(void *) * (DX_TAB_COLS + 64), cudaMemcpyHostToDevice, ctx->stream));
	*capacity = cap;
	*dest_cols = (void *const *) m->d_tab;
	*dest_counts = (unsigned long long *const *) (m->d_tab + DX_TAB_COLS);
	return CBGPU_OK;
}

extern "C" int
cbgpu_motion_direct_end(cbgpu_motion *m, const int64_t *dev_sent_counts, int64_t *sent_counts, cbgpu_rel **recv)
{
	int64_t		rows_sent_elsewhere = 0;

	cbgpu_ctx  *ctx = m->ctx;
	int		   *d_flag = (int *) (m->d_counts + (size_t) m->nranks * m->nranks);
	unsigned long long got = 0;
	int			rc;

	*recv = NULL;
	if (!m->direct_ok || m->dx_ncols < 1)
		return cb_fail(ctx, CBGPU_ERR_INVALID, "cbgpu_motion_direct_end without a begin%s", "", 0);
	/* stream order puts this after my kernel; its completion anywhere means every kernel is done */
	CB_CUDA(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream));
	CB_NCCL(ctx, ncclAllReduce(d_flag, d_flag, 1, ncclInt32, ncclMax, m->comm, ctx->stream));
	CB_CUDA(ctx, cudaMemcpyAsync(&got, m->win, sizeof(got), cudaMemcpyDeviceToHost, ctx->stream));
	/* the sender slice's own per-destination counts (a statistic) and the status word ride in the same round trip */
	if (dev_sent_counts && sent_counts)
		CB_CUDA(ctx, cudaMemcpyAsync(sent_counts, dev_sent_counts, sizeof(int64_t) * (size_t) m->nranks, cudaMemcpyDeviceToHost, ctx->stream));
	CB_CUDA(ctx, CB_STATUS_RIDE(ctx));
	if (ctx->trace_on)
		cb_trace_mark(ctx, "p2p:complete");
	CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	CB_STATUS_FETCHED(ctx);
	for (int d = 0; d < m->nranks && dev_sent_counts && sent_counts; d++)
		if (d != m->rank)
			rows_sent_elsewhere += sent_counts[d];
	if ((int64_t) got > m->dx_cap)
		return cb_fail(ctx, CBGPU_ERR_NOMEM, "Motion receive buffer overflowed (%s%lld rows): data skew beyond the reserved allowance", "", (long long) got);
	rc = cbgpu_rel_create(ctx, (int64_t) got, m->dx_ncols, m->dx_types, m->dx_dscales, recv);
	if (rc)
		return rc;
	if (got > 0 && got <= 65536)
	{
		/* a few rows (partial aggregate states, top-N candidates): one launch for all columns */
		CopyCols	cc;

		for (int c = 0; c < m->dx_ncols; c++)
		{
			cc.dst[c] = (*recv)->data[c];
			cc.src[c] = m->win + m->dx_off[c];
			cc.bytes[c] = (size_t) got * (size_t) cb_type_w(m->dx_types[c]);
		}
		k_copy_cols<<<m->dx_ncols, 256, 0, ctx->stream>>>(cc);
		CB_LAUNCHED(ctx, "k_copy_cols");
	}
	else
		for (int c = 0; c < m->dx_ncols && got > 0; c++)
			CB_CUDA(ctx, cudaMemcpyAsync((*recv)->data[c], m->win + m->dx_off[c], (size_t) got * (size_t) cb_type_w(m->dx_types[c]),
										 cudaMemcpyDeviceToDevice, ctx->stream));
	if (ctx->trace_on)
		cb_trace_mark(ctx, "p2p:copy-out");
	{
		int64_t		roww = 0;

		for (int c = 0; c < m->dx_ncols; c++)
			roww += cb_type_w(m->dx_types[c]);
		m->direct_bytes += rows_sent_elsewhere * roww;
	}
	m->direct_exchanges++;
	m->dx_ncols = 0;
	return CBGPU_OK;
}

/* direct Gather (a handful of rows per sender: final aggregates, top-N candidates): every sender
 * stores its row count and its rows into ITS slot of the root's Gather buffer, one 4-byte all-reduce
 * says "all stored" (and carries "somebody's rows did not fit": then everyone takes the NCCL path),
 * the root assembles.  Two buffers alternate, so the all-reduce of Gather k + 1 is also the proof that
 * the root has emptied the buffer Gather k + 2 will overwrite. */
struct GatherPut
{
	long long  *hdr;			/* root's count slot for this sender                                  */
	long long	nrows;
	CopyCols	cc;
	int			ncols;
};

__global__ void
k_gather_put(GatherPut g)
{
	if (blockIdx.x == 0 && threadIdx.x == 0)
		*g.hdr = g.nrows;
	if ((int) blockIdx.x < g.ncols)
	{
		const size_t n = g.cc.bytes[blockIdx.x];
		unsigned char *d = (unsigned char *) g.cc.dst[blockIdx.x];
		const unsigned char *s = (const unsigned char *) g.cc.src[blockIdx.x];

		for (size_t i = threadIdx.x; i < n; i += blockDim.x)
			d[i] = s[i];
	}
}

/* returns 1 when done directly, 0 when the caller must take the staged path, < 0 on error */
static int
gather_direct(cbgpu_motion *m, int root, cbgpu_rel *send, int64_t nrows, cbgpu_rel **recv)
{
	cbgpu_ctx  *ctx = m->ctx;
	const int	n = m->nranks;
	const size_t slot = ((GX_BYTES - GX_HDR) / (size_t) n) / DX_ALIGN * DX_ALIGN;
	const size_t buf = m->win_bytes - 2 * GX_BYTES + (size_t) (m->gather_seq & 1) * GX_BYTES;
	int		   *d_flag = (int *) (m->d_counts + (size_t) n * n);
	size_t		off[CBP_MAX_OUT];
	size_t		need = 0;
	int			fits = 1;
	int			h_flag = 0;
	long long	h_cnt[64];

	if (!m->direct_ok || n > 64 || send->ncols > CBP_MAX_OUT)
		return 0;
	for (int c = 0; c < send->ncols; c++)
	{
		if (send->nulls[c])
			fits = 0;			/* a null map may exist on one rank only: the decision rides on the all-reduce */
		off[c] = need;
		need += ((size_t) nrows * (size_t) cb_type_w(send->types[c]) + DX_ALIGN - 1) / DX_ALIGN * DX_ALIGN;
	}
	if (need > slot)
		fits = 0;
	m->gather_seq++;
	if (fits)
	{
		GatherPut	g;
		char	   *base = m->peer_win[root] + buf;

		memset(&g, 0, sizeof(g));
		g.hdr = (long long *) base + m->rank;
		g.nrows = nrows;
		g.ncols = send->ncols;
		for (int c = 0; c < send->ncols; c++)
		{
			g.cc.dst[c] = base + GX_HDR + (size_t) m->rank * slot + off[c];
			g.cc.src[c] = send->data[c];
			g.cc.bytes[c] = (size_t) nrows * (size_t) cb_type_w(send->types[c]);
		}
		k_gather_put<<<send->ncols > 0 ? send->ncols : 1, 256, 0, ctx->stream>>>(g);
		CB_LAUNCHED(ctx, "k_gather_put");
	}
	h_flag = fits ? 0 : 1;
	CB_CUDA(ctx, cudaMemcpyAsync(d_flag, &h_flag, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
	CB_NCCL(ctx, ncclAllReduce(d_flag, d_flag, 1, ncclInt32, ncclMax, m->comm, ctx->stream));
	CB_CUDA(ctx, cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	if (m->rank == root)
		CB_CUDA(ctx, cudaMemcpyAsync(h_cnt, m->win + buf, sizeof(long long) * (size_t) n, cudaMemcpyDeviceToHost, ctx->stream));
	if (ctx->trace_on)
		cb_trace_mark(ctx, "p2p:gather");
	CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	if (h_flag)
		return 0;				/* somebody's rows were too many for a slot: nothing was consumed, go staged */
	{
		int64_t		total = 0;
		int			rc;

		if (m->rank == root)
			for (int s = 0; s < n; s++)
				total += h_cnt[s];
		rc = make_recv(m, send, total, recv);
		if (rc)
			return -rc;
		if (m->rank == root && total > 0)
		{
			int64_t		at = 0;

			for (int s = 0; s < n; s++)
			{
				CopyCols	cc;

				if (h_cnt[s] == 0)
					continue;
				for (int c = 0; c < send->ncols; c++)
				{
					const size_t w = (size_t) cb_type_w(send->types[c]);
					/* every sender laid its columns out with ITS row count */
					size_t		o = 0;

					for (int cc2 = 0; cc2 < c; cc2++)
						o += ((size_t) h_cnt[s] * (size_t) cb_type_w(send->types[cc2]) + DX_ALIGN - 1) / DX_ALIGN * DX_ALIGN;
					cc.dst[c] = (char *) (*recv)->data[c] + (size_t) at * w;
					cc.src[c] = m->win + buf + GX_HDR + (size_t) s * slot + o;
					cc.bytes[c] = (size_t) h_cnt[s] * w;
				}
				k_copy_cols<<<send->ncols, 256, 0, ctx->stream>>>(cc);
				CB_LAUNCHED(ctx, "k_copy_cols");
				at += h_cnt[s];
			}
		}
		if (m->rank != root)
		{
			int64_t		roww = 0;

			for (int c = 0; c < send->ncols; c++)
				roww += cb_type_w(send->types[c]);
			m->direct_bytes += nrows * roww;
		}
	}
	m->direct_exchanges++;
	return 1;
}

extern "C" int
cbgpu_motion_gather(cbgpu_motion *m, int root, cbgpu_rel *send, int64_t nrows, cbgpu_rel **recv)
{
	int			n = m->nranks;
	int64_t    *matrix = (int64_t *) calloc((size_t) n * n, sizeof(int64_t));
	int64_t		mine[64], send_off[64], send_cnt[64], recv_off[64], recv_cnt[64];
	int64_t		total = 0;
	int			rc;

	if (n > 64)
		return cb_fail(m->ctx, CBGPU_ERR_UNSUPPORTED, "more than 64 segments%s", "", 0);
	rc = gather_direct(m, root, send, nrows, recv);
	if (rc != 0)
	{
		free(matrix);
		return rc > 0 ? CBGPU_OK : -rc;
	}
	for (int d = 0; d < n; d++)
		mine[d] = d == root ? nrows : 0;
	rc = exchange_counts(m, send, mine, matrix);
	if (rc == CBGPU_OK)
	{
		for (int s = 0; s < n; s++)
		{
			send_off[s] = 0;
			send_cnt[s] = mine[s];
			recv_off[s] = total;
			recv_cnt[s] = matrix[(size_t) s * n + m->rank];
			total += recv_cnt[s];
		}
		rc = make_recv(m, send, total, recv);
	}
	if (rc == CBGPU_OK)
		rc = exchange_rows(m, send, *recv, send_off, send_cnt, recv_off, recv_cnt);
	free(matrix);
	return rc;
}

extern "C" int
cbgpu_motion_broadcast(cbgpu_motion *m, cbgpu_rel *send, int64_t nrows, cbgpu_rel **recv)
{
	int			n = m->nranks;
	int64_t    *matrix = (int64_t *) calloc((size_t) n * n, sizeof(int64_t));
	int64_t		mine[64], send_off[64], send_cnt[64], recv_off[64], recv_cnt[64];
	int64_t		total = 0;
	int			rc;

	if (n > 64)
		return cb_fail(m->ctx, CBGPU_ERR_UNSUPPORTED, "more than 64 segments%s", "", 0);
	for (int d = 0; d < n; d++)
		mine[d] = nrows;
	rc = exchange_counts(m, send, mine, matrix);
	if (rc == CBGPU_OK)
	{
		for (int s = 0; s < n; s++)
		{
			send_off[s] = 0;
			send_cnt[s] = nrows;
			recv_off[s] = total;
			recv_cnt[s] = matrix[(size_t) s * n + m->rank];
			total += recv_cnt[s];
		}
		rc = make_recv(m, send, total, recv);
	}
	if (rc == CBGPU_OK)
		rc = exchange_rows(m, send, *recv, send_off, send_cnt, recv_off, recv_cnt);
	free(matrix);
	return rc;
}
