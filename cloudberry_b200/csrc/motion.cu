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

#include "../../include/cb_chan.h"

/* control block at the start of every window: the device-side signalling of the direct exchanges.
 * Slot [s] of each array is written by rank s only (remote stores over NVLink / peer memory). */
struct DxCtl
{
	unsigned long long counter;	/* rows reserved in this window by the senders of the exchange in flight   */
	unsigned long long pad[31];
	unsigned long long done[64];	/* done[s] = epoch << 16 | flags: rank s's rows of that exchange are stored */
	unsigned long long nmask[64];	/* columns for which rank s stored NULL bytes                              */
	unsigned long long ready[64];	/* written INTO rank s's window by ... see k_dx_release: ready[r] in MY
									 * window = last exchange rank r has finished taking delivery of         */
};

struct DxPeers
{
	DxCtl	   *ctl[64];
};

/* what k_dx_complete leaves in pinned host memory for the one round trip of an exchange */
struct DxResult
{
	unsigned long long got;		/* my window's row counter                                                  */
	unsigned long long flags;	/* union of every sender's flags (CBGPU_DX_*), DX_TIMEOUT if one never came */
	unsigned long long nmask;
	long long	cnt[64];		/* Gather: rows in every sender's slot                                      */
	long long	sent[64];		/* the sink's own per-destination counts                                    */
};

#define DX_TIMEOUT 0x8000u
#define DX_CTL_BYTES 4096
/* after the control block: the arena of the host packet channels (include/cb_chan.h; cbgpu_motion_chan_mem), then the
 * direct exchanges' data area */
#define CH_REGION_BYTES ((size_t) 8 << 20)
#define DX_DATA0 (DX_CTL_BYTES + CH_REGION_BYTES)
#define DX_TAB_COLS (64 * CBP_MAX_OUT)
#define DX_TAB_WORDS (2 * DX_TAB_COLS + 64)
#define DX_TAB_SLOTS 4
#define DX_ALIGN 256
/* the tail of every window: two alternating Gather / Broadcast buffers; each holds [64 row counts] + one
 * payload slot per sender */
#define GX_BYTES ((size_t) 32 << 20)
#define GX_HDR 1024

struct cbgpu_motion
{
	cbgpu_ctx  *ctx;
	ncclComm_t	comm;			/* NULL: windows only (cbgpu_motion_create_boot)                      */
	cbgpu_allgather_fn boot;	/* set-up / tear-down all-gather                                      */
	void	   *boot_arg;
	int			rank;
	int			nranks;
	long long  *d_counts;		/* [(nranks + 2)^2] scratch for the count exchange                    */
	int64_t		bytes_sent;		/* payload bytes this rank handed to NCCL (diagnostics / bench)       */
	int64_t		exchanges;
	/* peer-memory window: one cudaMalloc'ed arena per rank, mapped into every other rank's process
	 * (CUDA IPC), so a sender slice's PARTITION sink stores rows straight into the receiver's HBM
	 * over NVLink - no staging buffer, no payload through NCCL (direct Redistribute, below) */
	char	   *win;
	size_t		win_bytes;		/* the smallest window of all ranks                                   */
	char	   *peer_win[64];	/* peer_win[rank] == win                                              */
	bool		direct_ok;
	void	  **d_tab;			/* DX_TAB_SLOTS device tables handed to the sink: [DX_TAB_COLS] column bases,
								 * [DX_TAB_COLS] NULL byte bases, [64] counters                       */
	void	  **h_tab;			/* pinned host copies                                                 */
	int			tab_slot;
	int32_t    *d_flags;		/* the sink ORs CBGPU_DX_OVERFLOW in here                             */
	DxResult   *h_res;			/* pinned; k_dx_complete writes it                                    */
	DxPeers		peers;
	unsigned long long epoch;	/* direct exchanges so far (every rank runs the same sequence)        */
	unsigned long long timeout_ns;
	/* the direct exchange in flight */
	int32_t		dx_ncols;
	int32_t		dx_types[CBP_MAX_OUT];
	int32_t		dx_dscales[CBP_MAX_OUT];
	int64_t		dx_cap;
	size_t		dx_off[CBP_MAX_OUT];
	size_t		dx_noff[CBP_MAX_OUT];
	int64_t		direct_bytes;	/* payload bytes stored into peers' windows (diagnostics / bench)      */
	int64_t		direct_exchanges;
	int64_t		host_syncs;
	int64_t		collectives;
	cudaStream_t chan_stream;	/* the packet channels' copies: never queued behind the executor's kernels */
};

#define CB_NCCL(ctx, call) \
	do { \
		ncclResult_t r__ = (call); \
		if (r__ != ncclSuccess) \
		{ \
			snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: %s: %s", __FILE__, __LINE__, #call, ncclGetErrorString(r__)); \
			return CBGPU_ERR_CUDA; \
		} \
	} while (0)

#define NEED_NCCL(m, what) \
	do { \
		if (!(m)->comm) \
			return cb_fail((m)->ctx, CBGPU_ERR_UNSUPPORTED, "%s needs the NCCL transport, this interconnect has peer-memory windows only", what, 0); \
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

/* the set-up all-gather over the communicator itself */
static int
nccl_boot_allgather(void *arg, const void *mine, void *all, size_t bytes)
{
	cbgpu_motion *m = (cbgpu_motion *) arg;
	cbgpu_ctx  *ctx = m->ctx;
	char	   *d = NULL;
	const size_t n = (size_t) m->nranks;

	CB_CUDA(ctx, cudaMalloc(&d, bytes * (n + 1)));
	CB_CUDA(ctx, cudaMemcpyAsync(d + bytes * n, mine, bytes, cudaMemcpyHostToDevice, ctx->stream));
	CB_NCCL(ctx, ncclAllGather(d + bytes * n, d, bytes, ncclInt8, m->comm, ctx->stream));
	CB_CUDA(ctx, cudaMemcpyAsync(all, d, bytes * n, cudaMemcpyDeviceToHost, ctx->stream));
	CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	cudaFree(d);
	return 0;
}

static int
motion_create_common(cbgpu_ctx *ctx, int rank, int nranks, const void *unique_id128, cbgpu_allgather_fn boot, void *arg,
					 cbgpu_motion **out)
{
	cbgpu_motion *m = (cbgpu_motion *) calloc(1, sizeof(cbgpu_motion));
	const char *env = getenv("CBGPU_MOTION_TIMEOUT_MS");
	int			rc;

	*out = NULL;
	if (!m)
		return CBGPU_ERR_NOMEM;
	if (nranks < 1 || rank < 0 || rank >= nranks)
	{
		free(m);
		return cb_fail(ctx, CBGPU_ERR_INVALID, "interconnect rank %s%lld out of range", "", rank);
	}
	m->ctx = ctx;
	m->rank = rank;
	m->nranks = nranks;
	m->timeout_ns = (unsigned long long) (env && atoll(env) > 0 ? atoll(env) : 30000) * 1000000ull;
	CB_CUDA(ctx, cudaSetDevice(ctx->device));
	if (unique_id128)
	{
		ncclUniqueId id;

		memcpy(&id, unique_id128, 128);
		CB_NCCL(ctx, ncclCommInitRank(&m->comm, nranks, id, rank));
		m->boot = nccl_boot_allgather;
		m->boot_arg = m;
	}
	else
	{
		m->boot = boot;
		m->boot_arg = arg;
	}
	CB_CUDA(ctx, cudaMalloc(&m->d_counts, sizeof(long long) * (size_t) (nranks + 2) * (size_t) (nranks + 2)));
	*out = m;
	rc = motion_window_setup(m);
	if (rc == CBGPU_OK && !m->comm && !m->direct_ok)
		rc = cb_fail(ctx, CBGPU_ERR_UNSUPPORTED, "no peer-memory windows between the segments%s (CUDA IPC / P2P unavailable) and no NCCL transport", "", 0);
	return rc;
}

extern "C" int
cbgpu_motion_create(cbgpu_ctx *ctx, int rank, int nranks, const void *unique_id128, cbgpu_motion **out)
{
	if (!unique_id128)
		return cb_fail(ctx, CBGPU_ERR_INVALID, "cbgpu_motion_create without a rendezvous token%s", "", 0);
	return motion_create_common(ctx, rank, nranks, unique_id128, NULL, NULL, out);
}

extern "C" int
cbgpu_motion_create_boot(cbgpu_ctx *ctx, int rank, int nranks, cbgpu_allgather_fn allgather, void *arg, cbgpu_motion **out)
{
	if (!allgather)
		return cb_fail(ctx, CBGPU_ERR_INVALID, "cbgpu_motion_create_boot without an all-gather callback%s", "", 0);
	return motion_create_common(ctx, rank, nranks, NULL, allgather, arg, out);
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

/* allocate this rank's window, swap IPC handles through the set-up all-gather, map every peer's.
 * Any rank failing at any step turns the direct path off on ALL ranks (the decision is gathered too):
 * Redistribute then takes the staged NCCL path - both are device paths. */
static int
motion_window_setup(cbgpu_motion *m)
{
	cbgpu_ctx  *ctx = m->ctx;
	const int	n = m->nranks;
	WinHello   *h_all = (WinHello *) calloc((size_t) n, sizeof(WinHello));
	int		   *ok_all = (int *) calloc((size_t) n, sizeof(int));
	WinHello	mine;
	size_t		want = (size_t) 16384 << 20;	/* of 180 GB: room for a Motion of ~500 M narrow rows per receiver */
	const char *env = getenv("CBGPU_MOTION_WINDOW_MB");
	int			ok = 1;
	int			all_ok = 1;

	m->direct_ok = false;
	if (!h_all || !ok_all)
		return CBGPU_ERR_NOMEM;
	if (n > 64 || (m->comm && getenv("CBGPU_MOTION") && strcmp(getenv("CBGPU_MOTION"), "nccl") == 0))
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
		else
		{
			/* control block, and the NULL byte areas of every exchange start out zero (dx_release keeps them so) */
			CB_CUDA(ctx, cudaMemsetAsync(m->win, 0, want, ctx->stream));
			CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		}
	}
	mine.bytes = want;
	mine.ok = ok;
	if (m->boot(m->boot_arg, &mine, h_all, sizeof(WinHello)) != 0)
	{
		free(h_all);
		free(ok_all);
		return m->comm ? CBGPU_ERR_CUDA : cb_fail(ctx, CBGPU_ERR_PEER, "interconnect set-up: the caller's all-gather failed%s", "", 0);
	}
	m->collectives++;
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
	/* everyone or no one; this second all-gather is also the barrier "every window is zeroed and mapped" */
	if (m->boot(m->boot_arg, &ok, ok_all, sizeof(int)) != 0)
		all_ok = 0;
	m->collectives++;
	for (int p = 0; p < n; p++)
		if (!ok_all[p])
			all_ok = 0;
	free(h_all);
	free(ok_all);
	if (all_ok)
	{
		CB_CUDA(ctx, cudaMalloc(&m->d_tab, sizeof(void *) * DX_TAB_WORDS * DX_TAB_SLOTS));
		CB_CUDA(ctx, cudaMallocHost(&m->h_tab, sizeof(void *) * DX_TAB_WORDS * DX_TAB_SLOTS));
		CB_CUDA(ctx, cudaMalloc(&m->d_flags, sizeof(int32_t)));
		CB_CUDA(ctx, cudaMemsetAsync(m->d_flags, 0, sizeof(int32_t), ctx->stream));
		CB_CUDA(ctx, cudaMallocHost(&m->h_res, sizeof(DxResult)));
		memset(m->h_res, 0, sizeof(DxResult));
		for (int p = 0; p < n; p++)
			m->peers.ctl[p] = (DxCtl *) m->peer_win[p];
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
	if (m->d_flags)
		cudaFree(m->d_flags);
	if (m->h_res)
		cudaFreeHost(m->h_res);
	if (m->chan_stream)
		cudaStreamDestroy(m->chan_stream);
	m->chan_stream = NULL;
	m->d_tab = m->h_tab = NULL;
	m->d_flags = NULL;
	m->h_res = NULL;
	m->direct_ok = false;
}

extern "C" void
cbgpu_motion_destroy(cbgpu_motion *m)
{
	if (!m)
		return;
	cudaSetDevice(m->ctx->device);
	cudaStreamSynchronize(m->ctx->stream);
	if (m->direct_ok)
	{
		/* nobody unmaps a window a peer may still be storing into, nobody frees one a peer has mapped:
		 * (everyone has drained its stream) -> unmap the peers -> (everyone has unmapped) -> free */
		int			one = 1;
		int		   *all = (int *) calloc((size_t) m->nranks, sizeof(int));

		if (all)
			m->boot(m->boot_arg, &one, all, sizeof(int));
		for (int p = 0; p < m->nranks && p < 64; p++)
			if (p != m->rank && m->peer_win[p])
			{
				cudaIpcCloseMemHandle(m->peer_win[p]);
				m->peer_win[p] = NULL;
			}
		if (all)
			m->boot(m->boot_arg, &one, all, sizeof(int));
		free(all);
	}
	motion_window_teardown(m);
	if (m->comm)
		ncclCommDestroy(m->comm);
	cudaFree(m->d_counts);
	free(m);
}

extern "C" void
cbgpu_motion_abort(cbgpu_motion *m)
{
	if (!m)
		return;
	cudaSetDevice(m->ctx->device);
	if (m->comm)
		ncclCommAbort(m->comm);		/* outstanding collectives return; nothing is waited for */
	motion_window_teardown(m);
	cudaFree(m->d_counts);
	cudaGetLastError();
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

extern "C" int64_t
cbgpu_motion_host_syncs(const cbgpu_motion *m)
{
	return m->host_syncs;
}

extern "C" int64_t
cbgpu_motion_collectives(const cbgpu_motion *m)
{
	return m->collectives;
}

/* every rank learns every rank's per-destination row counts: matrix[s * nranks + d].  The same
 * all-gather carries which columns of `send` have a NULL map on each rank: a map may exist on one
 * segment only (it follows the data), but sender and receiver must post the same transfers, so every
 * rank gives `send` (zeroed) maps for the union before the rows move.  It also carries "this rank
 * failed before the exchange" (mine[0] < 0): then nobody posts a transfer and all return CBGPU_ERR_PEER
 * - a failed segment must not leave the others waiting inside a collective. */
static int
exchange_counts(cbgpu_motion *m, cbgpu_rel *send, const int64_t *mine, int64_t *matrix)
{
	cbgpu_ctx  *ctx = m->ctx;
	const int	n = m->nranks;
	long long  *d_all = m->d_counts;
	long long  *d_mine = m->d_counts + (size_t) (n + 2) * n;
	long long	h_mine[66], h_all[66 * 64];
	unsigned long long mask = 0,
				all = 0;
	const bool	failed = mine[0] < 0;
	int			peer_failed = -1;

	for (int c = 0; send && c < send->ncols && c < 64; c++)
		if (send->nulls[c])
			mask |= 1ull << c;
	for (int d = 0; d < n; d++)
		h_mine[d] = failed ? 0 : mine[d];
	h_mine[n] = (long long) mask;
	h_mine[n + 1] = failed ? 1 : 0;
	CB_CUDA(ctx, cudaMemcpyAsync(d_mine, h_mine, sizeof(long long) * (size_t) (n + 2), cudaMemcpyHostToDevice, ctx->stream));
	CB_NCCL(ctx, ncclAllGather(d_mine, d_all, (size_t) (n + 2), ncclInt64, m->comm, ctx->stream));
	CB_CUDA(ctx, cudaMemcpyAsync(h_all, d_all, sizeof(long long) * (size_t) (n + 2) * n, cudaMemcpyDeviceToHost, ctx->stream));
	if (ctx->trace_on)
		cb_trace_mark(ctx, "nccl:counts");
	CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	m->host_syncs++;
	m->collectives++;
	for (int s = 0; s < n; s++)
	{
		for (int d = 0; d < n; d++)
			matrix[(size_t) s * n + d] = h_all[(size_t) s * (n + 2) + d];
		all |= (unsigned long long) h_all[(size_t) s * (n + 2) + n];
		if (h_all[(size_t) s * (n + 2) + n + 1])
			peer_failed = s;
	}
	if (peer_failed >= 0)
		return cb_fail(ctx, CBGPU_ERR_PEER, "Motion abandoned: segment %s%lld failed before the exchange", "", peer_failed);
	for (int c = 0; send && c < send->ncols && c < 64; c++)
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
cbgpu_motion_redistribute(cbgpu_motion *m, cbgpu_rel *send, const int64_t *counts, const int64_t *offsets, cbgpu_rel **recv)
{
	int			n = m->nranks;
	int64_t    *matrix;
	int64_t		send_off[64], send_cnt[64], recv_off[64], recv_cnt[64];
	int64_t		total = 0;
	int			rc;

	*recv = NULL;
	NEED_NCCL(m, "a staged Redistribute Motion");
	if (n > 64)
		return cb_fail(m->ctx, CBGPU_ERR_UNSUPPORTED, "more than 64 segments%s", "", 0);
	matrix = (int64_t *) calloc((size_t) n * n, sizeof(int64_t));
	if (!matrix)
		return CBGPU_ERR_NOMEM;
	rc = exchange_counts(m, send, counts, matrix);
	if (rc == CBGPU_OK)
	{
		for (int s = 0; s < n; s++)
		{
			send_off[s] = offsets[s];
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
 * direct exchanges over the peer-memory windows, framed by device-side signals.
 *
 * Every rank runs the same sequence of exchanges (the plan is the same on every segment), numbered by
 * `epoch`.  Exchange e on rank r:
 *
 *   wait-ready   k_dx_wait_ready: spin (ld.acquire.sys on r's OWN window) until ready[p] >= e - 1 for every
 *                p: every receiver has taken delivery of exchange e - 1 and cleared its row counter
 *   store        the sender slice's pipeline kernel (PARTITION sink, direct mode): destination =
 *                cdbhashreduce(keys); a CTA reserves a slice of the DESTINATION's buffer with one
 *                system-scope atomic per destination and run, and stores the rows there over NVLink
 *   complete     k_dx_complete: fence, then done[r] = e << 16 | flags into EVERY window (st.release.sys);
 *                then spin until done[s] of exchange e has arrived from every sender s; the row counter,
 *                the union of the flags and the NULL-column masks land in pinned host memory
 *   (host)       ONE stream synchronisation: the receiver learns its row count and sizes the relation
 *   take         copy the rows out of the window, re-zero the NULL byte areas that were used
 *   release      k_dx_release: counter = 0, then ready[r] = e into every window
 *
 * No collective call, no host round trip except the one the executor needs anyway (a relation's row
 * count lives on the host).  A rank that fails locally still runs `complete` with CBGPU_DX_ERROR so nobody
 * waits for it; a rank that disappears is noticed by the spin loops' time limit (CBGPU_MOTION_TIMEOUT_MS,
 * default 30 s) -> CBGPU_ERR_PEER.  Overflow of a destination (skew beyond what the window holds) is
 * flagged by the sink, seen by every rank in `complete`, and answered by redoing the Motion staged with
 * exactly sized buffers: never by failing the query (the reference sends tuple by tuple and cannot
 * overflow, cdbmotion.c:425).
 * --------------------------------------------------------------------------------------------- */
__device__ __forceinline__ unsigned long long
dx_ld_acquire(const unsigned long long *p)
{
	unsigned long long v;

	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void
dx_st_relaxed(unsigned long long *p, unsigned long long v)
{
	asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long
dx_now_ns(void)
{
	unsigned long long t;

	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

/* spin until *p >= need (compared on the epoch part when shift > 0); false on time-out */
__device__ __forceinline__ bool
dx_spin(const unsigned long long *p, unsigned long long need, int shift, unsigned long long timeout_ns, unsigned long long *seen)
{
	const unsigned long long t0 = dx_now_ns();
	unsigned	spins = 0;

	for (;;)
	{
		const unsigned long long v = dx_ld_acquire(p);

		if ((v >> shift) >= need)
		{
			*seen = v;
			return true;
		}
		if (++spins > 64)
		{
			__nanosleep(spins > 4096 ? 2000 : 100);
			if ((spins & 255) == 0 && dx_now_ns() - t0 > timeout_ns)
			{
				*seen = v;
				return false;
			}
		}
	}
}

__global__ void
k_dx_wait_ready(const DxCtl *me, int n, unsigned long long need, unsigned long long timeout_ns, int *status)
{
	unsigned long long v;

	if ((int) threadIdx.x < n && !dx_spin(&me->ready[threadIdx.x], need, 0, timeout_ns, &v))
		atomicExch(status, CBGPU_ERR_PEER);
}

__global__ void
k_dx_complete(DxPeers peers, int me, int n, unsigned long long epoch, unsigned flags_host, const int32_t *dflags,
			  const int *status, unsigned long long nmask, const long long *ghdr, const int64_t *dev_sent, DxResult *res,
			  unsigned long long timeout_ns)
{
	__shared__ unsigned s_flags;
	__shared__ unsigned long long s_nmask;
	const int	p = threadIdx.x;

	if (p == 0)
	{
		s_flags = 0;
		s_nmask = 0;
	}
	__syncthreads();
	if (p < n)
	{
		/* an error one of my kernels raised (overflow, ...) fails the exchange for everybody: my rows are not to be used */
		const unsigned f = flags_host | (dflags ? (unsigned) *dflags : 0u) | (*status != 0 ? (unsigned) CBGPU_DX_ERROR : 0u);
		unsigned long long v = 0;

		/* my rows (stored by the kernels before me on this stream) before my signal, everywhere */
		__threadfence_system();
		dx_st_relaxed(&peers.ctl[p]->nmask[me], nmask);
		__threadfence_system();
		dx_st_relaxed(&peers.ctl[p]->done[me], (epoch << 16) | (f & 0xffffu));
		/* sender p's signal for this exchange */
		if (!dx_spin(&peers.ctl[me]->done[p], epoch, 16, timeout_ns, &v))
			atomicOr(&s_flags, DX_TIMEOUT);
		else
		{
			atomicOr(&s_flags, (unsigned) (v & 0xffffu));
			atomicOr(&s_nmask, dx_ld_acquire(&peers.ctl[me]->nmask[p]));
		}
	}
	__syncthreads();
	if (p < 64)
	{
		res->cnt[p] = (ghdr && p < n) ? ((const volatile long long *) ghdr)[p] : 0;
		/* Redistribute: the sink's per-destination counts; Gather: every sender's NULL-column mask */
		res->sent[p] = (dev_sent && p < n) ? dev_sent[p] : (ghdr && p < n) ? ((const volatile long long *) ghdr)[64 + p] : 0;
	}
	if (p == 0)
	{
		res->got = dx_ld_acquire(&peers.ctl[me]->counter);
		res->flags = s_flags;
		res->nmask = s_nmask;
	}
	__threadfence_system();
}

__global__ void
k_dx_release(DxPeers peers, int me, int n, unsigned long long epoch)
{
	if (threadIdx.x == 0)
		peers.ctl[me]->counter = 0;
	__syncthreads();
	if ((int) threadIdx.x < n)
	{
		__threadfence_system();
		dx_st_relaxed(&peers.ctl[threadIdx.x]->ready[me], epoch);
	}
}

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

/* start exchange number ++epoch: the device waits until every peer has released the previous one */
static int
dx_open(cbgpu_motion *m)
{
	cbgpu_ctx  *ctx = m->ctx;

	m->epoch++;
	k_dx_wait_ready<<<1, 64, 0, ctx->stream>>>((const DxCtl *) m->win, m->nranks, m->epoch - 1, m->timeout_ns, ctx->d_status);
	CB_LAUNCHED(ctx, "k_dx_wait_ready");
	return CBGPU_OK;
}

/* signal + wait + the exchange's one host round trip */
static int
dx_complete(cbgpu_motion *m, unsigned flags, unsigned long long nmask, const int32_t *dflags, const long long *ghdr,
			const int64_t *dev_sent, const char *mark)
{
	cbgpu_ctx  *ctx = m->ctx;

	k_dx_complete<<<1, 64, 0, ctx->stream>>>(m->peers, m->rank, m->nranks, m->epoch, flags, dflags, ctx->d_status, nmask, ghdr, dev_sent,
											 m->h_res, m->timeout_ns);
	CB_LAUNCHED(ctx, "k_dx_complete");
	CB_CUDA(ctx, CB_STATUS_RIDE(ctx));
	if (ctx->trace_on)
		cb_trace_mark(ctx, mark);
	CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	CB_STATUS_FETCHED(ctx);
	m->host_syncs++;
	return CBGPU_OK;
}

static int
dx_release(cbgpu_motion *m)
{
	cbgpu_ctx  *ctx = m->ctx;

	k_dx_release<<<1, 64, 0, ctx->stream>>>(m->peers, m->rank, m->nranks, m->epoch);
	CB_LAUNCHED(ctx, "k_dx_release");
	return CBGPU_OK;
}

static int
dx_peer_failure(cbgpu_motion *m, unsigned long long flags, const char *what)
{
	if (flags & DX_TIMEOUT)
		return cb_fail(m->ctx, CBGPU_ERR_PEER, "%s: a segment did not signal within the interconnect's time limit", what, 0);
	return cb_fail(m->ctx, CBGPU_ERR_PEER, "%s abandoned: another segment failed", what, 0);
}

extern "C" int
cbgpu_motion_direct_begin(cbgpu_motion *m, int32_t ncols, const int32_t *types, const int32_t *dscales, cbgpu_direct_dest *dest)
{
	cbgpu_ctx  *ctx = m->ctx;
	const int	n = m->nranks;
	size_t		off = DX_DATA0;
	size_t		roww = 0;
	size_t		avail;
	int64_t		cap;
	void	  **h, **d;

	memset(dest, 0, sizeof(*dest));
	if (!m->direct_ok)
		return cb_fail(ctx, CBGPU_ERR_UNSUPPORTED, "direct Redistribute is not available%s (no peer-memory window)", "", 0);
	if (ncols < 1 || ncols > CBP_MAX_OUT)
		return cb_fail(ctx, CBGPU_ERR_INVALID, "direct Redistribute of %s%lld columns", "", ncols);
	if (m->dx_ncols)
		return cb_fail(ctx, CBGPU_ERR_INVALID, "cbgpu_motion_direct_begin inside an open exchange%s", "", 0);
	/* the same arithmetic on every rank: the whole window (less the control block and the Gather buffers)
	 * divided by the row width, one NULL byte per column and row included */
	for (int c = 0; c < ncols; c++)
		roww += (size_t) cb_type_w(types[c]) + 1;
	avail = m->win_bytes - DX_DATA0 - 2 * GX_BYTES - (size_t) (2 * ncols) * DX_ALIGN;
	cap = (int64_t) (avail / roww) & ~(int64_t) 255;
	if (cap < 4096)
		return cb_fail(ctx, CBGPU_ERR_UNSUPPORTED, "direct Redistribute: a row of %s%lld bytes does not fit the window (raise CBGPU_MOTION_WINDOW_MB)", "", (long long) roww);
	for (int c = 0; c < ncols; c++)
	{
		m->dx_types[c] = types[c];
		m->dx_dscales[c] = dscales ? dscales[c] : 0;
		m->dx_off[c] = off;
		off += (((size_t) cap * (size_t) cb_type_w(types[c])) + DX_ALIGN - 1) / DX_ALIGN * DX_ALIGN;
	}
	for (int c = 0; c < ncols; c++)
	{
		m->dx_noff[c] = off;
		off += ((size_t) cap + DX_ALIGN - 1) / DX_ALIGN * DX_ALIGN;
	}
	m->dx_ncols = ncols;
	m->dx_cap = cap;
	m->tab_slot = (m->tab_slot + 1) % DX_TAB_SLOTS;
	h = m->h_tab + (size_t) m->tab_slot * DX_TAB_WORDS;
	d = m->d_tab + (size_t) m->tab_slot * DX_TAB_WORDS;
	for (int p = 0; p < n; p++)
	{
		for (int c = 0; c < ncols; c++)
		{
			h[(size_t) p * ncols + c] = m->peer_win[p] + m->dx_off[c];
			h[DX_TAB_COLS + (size_t) p * ncols + c] = m->peer_win[p] + m->dx_noff[c];
		}
		h[2 * DX_TAB_COLS + p] = m->peer_win[p];	/* DxCtl.counter */
	}
	CB_CUDA(ctx, cudaMemcpyAsync(d, h, sizeof(void *) * DX_TAB_WORDS, cudaMemcpyHostToDevice, ctx->stream));
	CB_CUDA(ctx, cudaMemsetAsync(m->d_flags, 0, sizeof(int32_t), ctx->stream));
	dx_open(m);
	if (ctx->trace_on)
		cb_trace_mark(ctx, "p2p:open");
	dest->capacity = cap;
	dest->cols = (void *const *) d;
	dest->nulls = (uint8_t *const *) (d + DX_TAB_COLS);
	dest->counts = (unsigned long long *const *) (d + 2 * DX_TAB_COLS);
	dest->flags = m->d_flags;
	return CBGPU_OK;
}

extern "C" int
cbgpu_motion_direct_end(cbgpu_motion *m, int32_t local_flags, uint64_t local_nullmask, const int64_t *dev_sent_counts,
						int64_t *sent_counts, cbgpu_rel **recv, int32_t *outcome)
{
	cbgpu_ctx  *ctx = m->ctx;
	DxResult	res;
	int64_t		got,
				used;
	int64_t		rows_sent_elsewhere = 0;
	const int	ncols = m->dx_ncols;
	int			rc;

	*recv = NULL;
	*outcome = CBGPU_DX_DELIVERED;
	if (!m->direct_ok || ncols < 1)
		return cb_fail(ctx, CBGPU_ERR_INVALID, "cbgpu_motion_direct_end without a begin%s", "", 0);
	m->dx_ncols = 0;
	rc = dx_complete(m, (unsigned) local_flags, local_nullmask, m->d_flags, NULL, dev_sent_counts, "p2p:complete");
	if (rc)
		return rc;
	res = *m->h_res;
	got = (int64_t) res.got;
	used = got < m->dx_cap ? got : m->dx_cap;
	for (int d = 0; d < m->nranks; d++)
	{
		if (sent_counts)
			sent_counts[d] = res.sent[d];
		if (d != m->rank)
			rows_sent_elsewhere += res.sent[d];
	}
	/* NULL bytes a sender stored must be gone before the next exchange (senders without NULLs store none) */
	if (!(res.flags & (DX_TIMEOUT | CBGPU_DX_ERROR | CBGPU_DX_OVERFLOW | CBGPU_DX_VETO)))
	{
		rc = cbgpu_rel_create(ctx, got, ncols, m->dx_types, m->dx_dscales, recv);
		if (rc)
			return rc;
		for (int c = 0; c < ncols; c++)
			if ((res.nmask >> c) & 1)
			{
				rc = cbgpu_rel_add_nullmap(*recv, c);
				if (rc)
					return rc;
				if (got > 0)
					CB_CUDA(ctx, cudaMemcpyAsync((*recv)->nulls[c], m->win + m->dx_noff[c], (size_t) got, cudaMemcpyDeviceToDevice, ctx->stream));
			}
		if (got > 0 && got <= 65536)
		{
			/* a few rows (partial aggregate states, top-N candidates): one launch for all columns */
			CopyCols	cc;

			for (int c = 0; c < ncols; c++)
			{
				cc.dst[c] = (*recv)->data[c];
				cc.src[c] = m->win + m->dx_off[c];
				cc.bytes[c] = (size_t) got * (size_t) cb_type_w(m->dx_types[c]);
			}
			k_copy_cols<<<ncols, 256, 0, ctx->stream>>>(cc);
			CB_LAUNCHED(ctx, "k_copy_cols");
		}
		else
			for (int c = 0; c < ncols && got > 0; c++)
				CB_CUDA(ctx, cudaMemcpyAsync((*recv)->data[c], m->win + m->dx_off[c], (size_t) got * (size_t) cb_type_w(m->dx_types[c]),
											 cudaMemcpyDeviceToDevice, ctx->stream));
		if (ctx->trace_on)
			cb_trace_mark(ctx, "p2p:copy-out");
	}
	for (int c = 0; c < ncols && used > 0; c++)
		if ((res.nmask >> c) & 1)
			CB_CUDA(ctx, cudaMemsetAsync(m->win + m->dx_noff[c], 0, (size_t) used, ctx->stream));
	dx_release(m);
	if (res.flags & (DX_TIMEOUT | CBGPU_DX_ERROR))
		return dx_peer_failure(m, res.flags, "direct Redistribute");
	if (res.flags & (CBGPU_DX_OVERFLOW | CBGPU_DX_VETO))
	{
		*outcome = CBGPU_DX_RETRY;
		return CBGPU_OK;
	}
	{
		int64_t		roww = 0;

		for (int c = 0; c < ncols; c++)
			roww += cb_type_w(m->dx_types[c]);
		m->direct_bytes += rows_sent_elsewhere * roww;
	}
	m->direct_exchanges++;
	return CBGPU_OK;
}

extern "C" int
cbgpu_motion_abandon(cbgpu_motion *m, int staged)
{
	if (m->dx_ncols)
	{
		/* inside an open direct exchange: close it with the error flag */
		cbgpu_rel  *recv = NULL;
		int32_t		outcome;

		cbgpu_motion_direct_end(m, CBGPU_DX_ERROR, 0, NULL, NULL, &recv, &outcome);
		if (recv)
			cbgpu_rel_free(recv);
		return CBGPU_OK;
	}
	if (!staged && m->direct_ok)
	{
		dx_open(m);
		dx_complete(m, CBGPU_DX_ERROR, 0, NULL, NULL, NULL, "p2p:abandon");
		dx_release(m);
		return CBGPU_OK;
	}
	if (m->comm)
	{
		int64_t		mine[64];
		int64_t    *matrix = (int64_t *) calloc((size_t) m->nranks * m->nranks, sizeof(int64_t));

		memset(mine, 0, sizeof(mine));
		mine[0] = -1;
		if (matrix)
			exchange_counts(m, NULL, mine, matrix);
		free(matrix);
	}
	return CBGPU_OK;
}

/* direct Gather / Broadcast (a handful of rows per sender: final aggregates, top-N candidates): every
 * sender stores its row count and its rows into ITS slot of the Gather buffer of the root (Broadcast: of
 * every rank), `complete` says "all stored" and carries "somebody's rows did not fit" (CBGPU_DX_NOFIT:
 * then nothing is consumed and everyone takes the NCCL path), the receiver assembles.  Two buffers
 * alternate by epoch parity; `release` of exchange e is the proof that a buffer of exchange e - 1 is empty. */
struct GatherPut
{
	long long  *hdr[64];		/* the receivers' count slot for this sender (its NULL-column mask 64 slots on) */
	char	   *base[64];		/* ... and payload slot                                               */
	int			ndest;
	long long	nrows;
	long long	nullmask;
	const void *src[2 * CBP_MAX_OUT];	/* the columns, then the NULL maps of the columns that have one     */
	size_t		off[2 * CBP_MAX_OUT];
	size_t		bytes[2 * CBP_MAX_OUT];
	int			nparts;
};

__global__ void
k_gather_put(GatherPut g)
{
	const int	d = blockIdx.y;

	if (blockIdx.x == 0 && threadIdx.x == 0)
	{
		g.hdr[d][0] = g.nrows;
		g.hdr[d][64] = g.nullmask;
	}
	if ((int) blockIdx.x < g.nparts)
	{
		const size_t n = g.bytes[blockIdx.x];
		unsigned char *dst = (unsigned char *) g.base[d] + g.off[blockIdx.x];
		const unsigned char *s = (const unsigned char *) g.src[blockIdx.x];

		for (size_t i = threadIdx.x; i < n; i += blockDim.x)
			dst[i] = s[i];
	}
}

static size_t
gx_pad(size_t bytes)
{
	return (bytes + DX_ALIGN - 1) / DX_ALIGN * DX_ALIGN;
}

/* root >= 0: Gather to that rank; root < 0: Broadcast.  Returns 1 when done directly, 0 when the caller must
 * take the staged path (decided identically on every rank), < 0 on error.  A sender lays its slot out as
 * [column 0][column 1]...[NULL bytes of its nullable columns, in column order], every part padded to
 * DX_ALIGN; row count and NULL-column mask go into the buffer's header. */
static int
gather_direct(cbgpu_motion *m, int root, cbgpu_rel *send, int64_t nrows, cbgpu_rel **recv)
{
	cbgpu_ctx  *ctx = m->ctx;
	const int	n = m->nranks;
	const size_t slot = ((GX_BYTES - GX_HDR) / (size_t) n) / DX_ALIGN * DX_ALIGN;
	size_t		buf;
	size_t		need = 0;
	int			fits = 1;
	const bool	receiver = root < 0 || m->rank == root;
	GatherPut	g;
	DxResult	res;

	if (!m->direct_ok || n > 64 || send->ncols > CBP_MAX_OUT)
		return 0;
	memset(&g, 0, sizeof(g));
	for (int c = 0; c < send->ncols; c++)
	{
		g.src[g.nparts] = send->data[c];
		g.bytes[g.nparts] = (size_t) nrows * (size_t) cb_type_w(send->types[c]);
		g.off[g.nparts] = need;
		need += gx_pad(g.bytes[g.nparts]);
		g.nparts++;
	}
	for (int c = 0; c < send->ncols; c++)
		if (send->nulls[c])
		{
			g.nullmask |= 1ll << c;
			g.src[g.nparts] = send->nulls[c];
			g.bytes[g.nparts] = (size_t) nrows;
			g.off[g.nparts] = need;
			need += gx_pad(g.bytes[g.nparts]);
			g.nparts++;
		}
	if (need > slot)
		fits = 0;
	dx_open(m);
	buf = m->win_bytes - 2 * GX_BYTES + (size_t) (m->epoch & 1) * GX_BYTES;
	if (fits)
	{
		for (int p = 0; p < n; p++)
			if (root < 0 || p == root)
			{
				g.hdr[g.ndest] = (long long *) (m->peer_win[p] + buf) + m->rank;
				g.base[g.ndest] = m->peer_win[p] + buf + GX_HDR + (size_t) m->rank * slot;
				g.ndest++;
			}
		g.nrows = nrows;
		k_gather_put<<<dim3(g.nparts > 0 ? g.nparts : 1, g.ndest), 256, 0, ctx->stream>>>(g);
		CB_LAUNCHED(ctx, "k_gather_put");
	}
	if (dx_complete(m, fits ? 0u : CBGPU_DX_NOFIT, 0, NULL, receiver ? (const long long *) (m->win + buf) : NULL, NULL, "p2p:gather") != CBGPU_OK)
		return -CBGPU_ERR_CUDA;
	res = *m->h_res;
	if (res.flags & (DX_TIMEOUT | CBGPU_DX_ERROR))
	{
		dx_release(m);
		return -dx_peer_failure(m, res.flags, root < 0 ? "Broadcast Motion" : "Gather Motion");
	}
	if (res.flags & CBGPU_DX_NOFIT)
	{
		dx_release(m);
		return 0;				/* somebody's rows were too many for a slot: nothing was consumed, go staged */
	}
	{
		int64_t		total = 0;
		unsigned long long anynull = 0;
		int			rc;

		if (receiver)
			for (int s = 0; s < n; s++)
			{
				total += res.cnt[s];
				if (res.cnt[s] > 0)
					anynull |= (unsigned long long) res.sent[s];
			}
		rc = cbgpu_rel_create(ctx, total, send->ncols, send->types, send->dscales, recv);
		for (int c = 0; c < send->ncols && rc == CBGPU_OK; c++)
		{
			if ((anynull >> c) & 1)
				rc = cbgpu_rel_add_nullmap(*recv, c);	/* zero-filled: senders without a map for it sent no NULLs */
			if (send->dict_hash[c])
				cbgpu_rel_share_dict_hash(*recv, c, send, c);
		}
		if (rc)
		{
			dx_release(m);
			return -rc;
		}
		if (receiver && total > 0)
		{
			int64_t		at = 0;

			for (int s = 0; s < n; s++)
			{
				CopyCols	cc;
				size_t		o = 0;
				int			parts = 0;

				if (res.cnt[s] == 0)
					continue;
				/* every sender laid its slot out with ITS row count and ITS NULL maps */
				for (int c = 0; c < send->ncols; c++)
				{
					const size_t w = (size_t) cb_type_w(send->types[c]);

					cc.dst[parts] = (char *) (*recv)->data[c] + (size_t) at * w;
					cc.src[parts] = m->win + buf + GX_HDR + (size_t) s * slot + o;
					cc.bytes[parts] = (size_t) res.cnt[s] * w;
					o += gx_pad(cc.bytes[parts]);
					parts++;
				}
				k_copy_cols<<<parts, 256, 0, ctx->stream>>>(cc);
				CB_LAUNCHED(ctx, "k_copy_cols");
				parts = 0;
				for (int c = 0; c < send->ncols; c++)
					if (((unsigned long long) res.sent[s] >> c) & 1)
					{
						cc.dst[parts] = (char *) (*recv)->nulls[c] + (size_t) at;
						cc.src[parts] = m->win + buf + GX_HDR + (size_t) s * slot + o;
						cc.bytes[parts] = (size_t) res.cnt[s];
						o += gx_pad(cc.bytes[parts]);
						parts++;
					}
				if (parts)
				{
					k_copy_cols<<<parts, 256, 0, ctx->stream>>>(cc);
					CB_LAUNCHED(ctx, "k_copy_cols");
				}
				at += res.cnt[s];
			}
		}
		{
			int64_t		roww = 0;

			for (int c = 0; c < send->ncols; c++)
				roww += cb_type_w(send->types[c]);
			m->direct_bytes += nrows * roww * (root < 0 ? n - 1 : (m->rank != root ? 1 : 0));
		}
	}
	dx_release(m);
	m->direct_exchanges++;
	return 1;
}

extern "C" int
cbgpu_motion_gather(cbgpu_motion *m, int root, cbgpu_rel *send, int64_t nrows, cbgpu_rel **recv)
{
	int			n = m->nranks;
	int64_t    *matrix;
	int64_t		mine[64] = {0}, send_off[64], send_cnt[64], recv_off[64], recv_cnt[64];
	int64_t		total = 0;
	int			rc;

	*recv = NULL;
	if (n > 64 || root < 0 || root >= n)
		return cb_fail(m->ctx, CBGPU_ERR_UNSUPPORTED, "Gather Motion to segment %s%lld", "", root);
	rc = gather_direct(m, root, send, nrows, recv);
	if (rc != 0)
		return rc > 0 ? CBGPU_OK : -rc;
	NEED_NCCL(m, "a Gather Motion too large for the window's Gather slots");
	matrix = (int64_t *) calloc((size_t) n * n, sizeof(int64_t));
	if (!matrix)
		return CBGPU_ERR_NOMEM;
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
	int64_t    *matrix;
	int64_t		mine[64] = {0}, send_off[64], send_cnt[64], recv_off[64], recv_cnt[64];
	int64_t		total = 0;
	int			rc;

	*recv = NULL;
	if (n > 64)
		return cb_fail(m->ctx, CBGPU_ERR_UNSUPPORTED, "more than 64 segments%s", "", 0);
	rc = gather_direct(m, -1, send, nrows, recv);
	if (rc != 0)
		return rc > 0 ? CBGPU_OK : -rc;
	NEED_NCCL(m, "a Broadcast Motion too large for the window's Gather slots");
	matrix = (int64_t *) calloc((size_t) n * n, sizeof(int64_t));
	if (!matrix)
		return CBGPU_ERR_NOMEM;
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

/* ---------------------------------------------------------------------------------------------
 * host packet channels over the windows (include/cb_chan.h): the arena is CH_REGION_BYTES of every window, a put is a
 * host -> PEER-device copy (the copy engine stores over NVLink into the peer's HBM; copies on one stream land in
 * order, so a packet is complete before the counter that publishes it), a get a device -> host copy of this rank's
 * own window.  What integration/cbgpu_ic_layer.c moves tuple chunks with between QE processes.
 * --------------------------------------------------------------------------------------------- */
static int
win_chan_put(void *arg, int peer, size_t off, const void *src, size_t len)
{
	cbgpu_motion *m = (cbgpu_motion *) arg;

	if (peer < 0 || peer >= m->nranks || off + len > CH_REGION_BYTES)
		return -1;
	/* pageable source: the call returns once the bytes are staged, the caller may reuse its buffer */
	return cudaMemcpyAsync(m->peer_win[peer] + DX_CTL_BYTES + off, src, len, cudaMemcpyDefault, m->chan_stream) == cudaSuccess ? 0 : -1;
}

static int
win_chan_get(void *arg, size_t off, void *dst, size_t len)
{
	cbgpu_motion *m = (cbgpu_motion *) arg;

	if (off + len > CH_REGION_BYTES)
		return -1;
	if (cudaMemcpyAsync(dst, m->win + DX_CTL_BYTES + off, len, cudaMemcpyDefault, m->chan_stream) != cudaSuccess)
		return -1;
	return cudaStreamSynchronize(m->chan_stream) == cudaSuccess ? 0 : -1;
}

extern "C" int
cbgpu_motion_chan_mem(cbgpu_motion *m, CbChanMem *mem, size_t *arena_bytes)
{
	if (!m->direct_ok)
		return cb_fail(m->ctx, CBGPU_ERR_UNSUPPORTED, "packet channels need the peer-memory windows%s", "", 0);
	if (!m->chan_stream)
		CB_CUDA(m->ctx, cudaStreamCreateWithFlags(&m->chan_stream, cudaStreamNonBlocking));
	mem->arg = m;
	mem->put = win_chan_put;
	mem->get = win_chan_get;
	*arena_bytes = CH_REGION_BYTES;
	return CBGPU_OK;
}
