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

struct cbgpu_motion
{
	cbgpu_ctx  *ctx;
	ncclComm_t	comm;
	int			rank;
	int			nranks;
	long long  *d_counts;		/* [nranks * nranks] scratch for the count exchange                   */
	int64_t		bytes_sent;		/* payload bytes this rank handed to NCCL (diagnostics / bench)       */
	int64_t		exchanges;
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
	CB_CUDA(ctx, cudaMalloc(&m->d_counts, sizeof(long long) * (size_t) nranks * (size_t) (nranks + 1)));
	*out = m;
	return CBGPU_OK;
}

extern "C" void
cbgpu_motion_destroy(cbgpu_motion *m)
{
	if (!m)
		return;
	cudaSetDevice(m->ctx->device);
	cudaStreamSynchronize(m->ctx->stream);
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

/* every rank learns every rank's per-destination row counts: matrix[s * nranks + d] */
static int
exchange_counts(cbgpu_motion *m, const int64_t *mine, int64_t *matrix)
{
	cbgpu_ctx  *ctx = m->ctx;
	int			n = m->nranks;
	long long  *d_mine = m->d_counts + (size_t) n * n;

	CB_CUDA(ctx, cudaMemcpyAsync(d_mine, mine, sizeof(long long) * n, cudaMemcpyHostToDevice, ctx->stream));
	CB_NCCL(ctx, ncclAllGather(d_mine, m->d_counts, (size_t) n, ncclInt64, m->comm, ctx->stream));
	CB_CUDA(ctx, cudaMemcpyAsync(matrix, m->d_counts, sizeof(long long) * (size_t) n * n, cudaMemcpyDeviceToHost, ctx->stream));
	CB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
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
	rc = exchange_counts(m, counts, matrix);
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
	for (int d = 0; d < n; d++)
		mine[d] = d == root ? nrows : 0;
	rc = exchange_counts(m, mine, matrix);
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
	rc = exchange_counts(m, mine, matrix);
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
