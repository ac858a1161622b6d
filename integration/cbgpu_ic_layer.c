/*
 * cbgpu_ic_layer.c - a MotionIPCLayer (include/cdb/ml_ipc.h:36-294) that moves the reference's tuple chunks between QE
 * processes over packet channels (include/cb_chan.h): rings in the receivers' arenas - GPU peer-memory windows written
 * over NVLink, or POSIX shared memory on one host - instead of UDP / TCP sockets.  SURVEY.md 8 row f3.
 *
 * It is a separate loadable module, like contrib/udp2 (contrib/udp2/ic_modules.c:87): _PG_init calls
 * RegisterIPCLayerImpl (cdb/motion/cdbmotion.c:1315).  That function rejects a second implementation of a
 * *registered* ic_type and a full table (MAX_NUMBER_TYPES = 4, cdbvars.h:294-299); the GUC picks an implementation by
 * its type_name string (SetCurrentMotionIPCLayer, cdbmotion.c:1259) and core never branches on the numeric value.  So on
 * a cluster that does not load udp2, this module takes that free slot: ic_type = INTERCONNECT_TYPE_UDP2, type_name =
 * "cbgpu", and `gp_interconnect_type = cbgpu` selects it.  No core patch.
 *
 * What the executor's Motion nodes see is unchanged: SendTuple / RecvTupleFrom (cdbmotion.c:425,549) serialise tuples to
 * chunks and hand them to this layer; unreplaced CPU operators on one segment exchange rows with GPU operators on another
 * through it (mixed plans), byte for byte what tupser.c produces.
 *
 * Endpoints.  Every process of the query (the QD and every QE of every gang) is one channel endpoint, numbered by its
 * place in the slice table, which all of them hold: rank(slice k, process i) = sum of the gang sizes of slices < k, + i.
 * A Motion's routes are the positions in the peer slice's primaryProcesses list, as in the reference
 * (createChunkTransportState, contrib/interconnect/ic_common.c:243).  Rendezvous needs no listener socket: an
 * endpoint's arena is named after (gp_session_id, gp_command_count, rank), which every process can compute; a peer that
 * is not there yet is waited for (Gp_interconnect_timeout).
 *
 * Packets.  One channel packet = CbIcPktHdr + whole tuple chunks (4-byte chunk headers, cdb/tupchunk.h) of ONE
 * (motion node, sender route).  A connection's chunks collect in its transmit buffer and go out when it is full, at
 * end of stream, and before the sender blocks in a receive of its own (so two slices can never wait on each other's
 * unflushed data).  Flow control is the channel's: a full ring at the receiver makes the send wait - while it waits it
 * keeps draining its OWN incoming rings into per-connection queues, the way the reference's senders keep polling acks.
 *
 * Type-checked against the reference's headers (tests/test_shim_compiles.py) and RUN against the reference's own
 * cdbmotion.c / tupser.c compiled where they lie (oracle/ref_motion.c, tests/test_ic_layer.py): tuples serialised by the
 * reference's SendTuple come out of the reference's RecvTupleFrom on another process bit for bit.
 */
#include "postgres.h"

#include "cdb/cdbgang.h"
#include "cdb/cdbinterconnect.h"
#include "cdb/cdbmotion.h"
#include "cdb/cdbvars.h"
#include "cdb/ml_ipc.h"
#include "cdb/tupchunk.h"
#include "cdb/tupchunklist.h"
#include "cdb/tupleremap.h"
#include "fmgr.h"
#include "miscadmin.h"
#include "nodes/execnodes.h"
#include "nodes/pg_list.h"
#include "utils/memutils.h"

#include "cb_chan.h"
#include "cbgpu.h"

#ifndef CBGPU_IC_NO_MODULE_MAGIC
PG_MODULE_MAGIC;
#endif

/* the socket interconnects' cancel check (contrib/interconnect/ic_common.h:82): inside a teardown nothing longjmps */
#define ML_CHECK_FOR_INTERRUPTS(teardownActive) \
	do { \
		if (!(teardownActive) && InterruptPending) \
			CHECK_FOR_INTERRUPTS(); \
	} while (0)

#define CBIC_MAGIC 0xCB1C
#define CBIC_FLAG_EOS 1			/* the packet ends this sender's stream (it carries the TC_END_OF_STREAM chunk)         */
#define CBIC_FLAG_STOP 2		/* receiver -> sender: no more tuples wanted (SendStopMessage)                          */
#define CBIC_SLOTS 8
#define CBIC_SLOT_BYTES 65536

typedef struct CbIcPktHdr
{
	uint16		magic;
	int16		motNodeID;
	int16		srcRoute;		/* the sender's position in its slice (STOP: the receiver's position in its slice)      */
	uint16		flags;
	uint32		nbytes;			/* chunk bytes that follow                                                           */
	uint32		seq;
} CbIcPktHdr;

typedef struct CbIcPkt			/* a received packet waiting in its connection's queue */
{
	struct CbIcPkt *next;
	uint32		nbytes;
	uint16		flags;
	unsigned char data[FLEXIBLE_ARRAY_MEMBER];
} CbIcPkt;

typedef struct CbIcConn
{
	int			rank;			/* channel endpoint of the peer process                                              */
	/* sending side */
	bool		still_active;	/* the receiver still wants tuples                                                   */
	bool		eos_sent;
	unsigned char *tx;			/* CbIcPktHdr + chunks                                                               */
	int			txused;
	uint32		txseq;
	/* receiving side */
	bool		eos_seen;
	bool		deregistered;
	CbIcPkt    *rx_head,
			   *rx_tail;
	CbIcPkt    *rx_held;		/* the packet whose chunks the caller is looking at (TupleChunkListItem.inplace points into it)
								 * until DirectPutRxBuffer hands it back                                             */
	TupleRemapper *remapper;
	int32		sent_record_typmod;
} CbIcConn;

typedef struct CbIcNode
{
	int16		motNodeID;
	bool		is_sender;
	int			nconns;
	CbIcConn   *conns;
	int			next_any;		/* RecvTupleChunkFromAny: where the fair scan of the queues resumes                  */
} CbIcNode;

typedef struct CbIcState
{
	MemoryContext cxt;
	cb_chan    *chan;
	cb_chan_shm *shm;
	int			rank,
				nranks;
	int			my_route;		/* my position in my own slice: the srcRoute of what I send                         */
	int			nnodes;
	CbIcNode   *nodes;
	unsigned char *rxbuf;
	int			max_packet;
} CbIcState;

static int	cbic_active_conns = 0;

/* ------------------------------------------------------------------------------------------
 * helpers
 * ------------------------------------------------------------------------------------------ */
static CbIcState *
cbic_state(ChunkTransportState *ts)
{
	if (ts == NULL || ts->implement_state == NULL)
		ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: no transport state")));
	if (!ts->activated)
		ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: transport state inactive")));
	return (CbIcState *) ts->implement_state;
}

static CbIcNode *
cbic_node(CbIcState *st, int motNodeID, bool sender)
{
	for (int i = 0; i < st->nnodes; i++)
		if (st->nodes[i].motNodeID == motNodeID && st->nodes[i].is_sender == sender)
			return &st->nodes[i];
	ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR),
					errmsg("cbgpu interconnect: this process is no %s of motion node %d", sender ? "sender" : "receiver", motNodeID)));
	return NULL;
}

/* file an incoming packet under its (receiving motion node, sender route); STOP packets act on the sending side */
static void
cbic_file_packet(CbIcState *st, const unsigned char *buf, int len, int from_rank)
{
	CbIcPktHdr	h;
	CbIcNode   *node;
	CbIcConn   *conn;
	CbIcPkt    *p;

	if (len < (int) sizeof(h))
		ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: short packet (%d bytes) from endpoint %d", len, from_rank)));
	memcpy(&h, buf, sizeof(h));
	if (h.magic != CBIC_MAGIC || (int) (sizeof(h) + h.nbytes) != len)
		ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: malformed packet from endpoint %d", from_rank)));
	if (h.flags & CBIC_FLAG_STOP)
	{
		/* a receiver of MY stream wants no more: the reference marks the connection !stillActive (handleStopMsgs) */
		for (int i = 0; i < st->nnodes; i++)
			if (st->nodes[i].is_sender && st->nodes[i].motNodeID == h.motNodeID && h.srcRoute >= 0 && h.srcRoute < st->nodes[i].nconns)
				st->nodes[i].conns[h.srcRoute].still_active = false;
		return;
	}
	node = cbic_node(st, h.motNodeID, false);
	if (h.srcRoute < 0 || h.srcRoute >= node->nconns || node->conns[h.srcRoute].rank != from_rank)
		ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR),
						errmsg("cbgpu interconnect: packet for motion node %d claims route %d, came from endpoint %d", h.motNodeID, h.srcRoute, from_rank)));
	conn = &node->conns[h.srcRoute];
	if (conn->deregistered)
		return;					/* nobody reads this stream any more */
	p = MemoryContextAlloc(st->cxt, offsetof(CbIcPkt, data) + h.nbytes);
	p->next = NULL;
	p->nbytes = h.nbytes;
	p->flags = h.flags;
	memcpy(p->data, buf + sizeof(h), h.nbytes);
	if (conn->rx_tail)
		conn->rx_tail->next = p;
	else
		conn->rx_head = p;
	conn->rx_tail = p;
}

/* take whatever has arrived on any ring (waiting up to wait_ms for the first packet); returns packets filed */
static int
cbic_drain(CbIcState *st, int wait_ms)
{
	int			n = 0;

	for (;;)
	{
		int			from = -1;
		int			len = cb_chan_recv(st->chan, -1, st->rxbuf, st->max_packet, &from, n == 0 ? wait_ms : 0);

		if (len < 0)
			ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: receive failed (%d)", len)));
		if (len == 0)
			return n;
		cbic_file_packet(st, st->rxbuf, len, from);
		n++;
	}
}

/* send one connection's transmit buffer (if it holds chunks or `flags` must travel) */
static void
cbic_flush(ChunkTransportState *ts, CbIcState *st, CbIcNode *node, CbIcConn *conn, uint16 flags)
{
	CbIcPktHdr	h;
	int64		waited = 0;

	if (conn->txused == (int) sizeof(CbIcPktHdr) && flags == 0)
		return;
	h.magic = CBIC_MAGIC;
	h.motNodeID = node->motNodeID;
	h.srcRoute = (int16) st->my_route;
	h.flags = flags;
	h.nbytes = (uint32) (conn->txused - (int) sizeof(CbIcPktHdr));
	h.seq = conn->txseq++;
	memcpy(conn->tx, &h, sizeof(h));
	for (;;)
	{
		int			rc = cb_chan_send(st->chan, conn->rank, conn->tx, conn->txused, 20);

		if (rc == 0)
			break;
		if (rc < 0)
			ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: send to endpoint %d failed", conn->rank)));
		/* the receiver's ring is full: its flow control.  Meanwhile keep my own rings moving (nobody may wait on me) and
		 * stay cancellable, as the reference's senders poll for acks and interrupts (ic_udpifc.c sendLoop) */
		ML_CHECK_FOR_INTERRUPTS(ts->teardownActive);
		cbic_drain(st, 0);
		waited += 20;
		if (Gp_interconnect_transmit_timeout > 0 && waited > (int64) Gp_interconnect_transmit_timeout * 1000)
			ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR),
							errmsg("cbgpu interconnect: endpoint %d accepted no packet for %d s", conn->rank, Gp_interconnect_transmit_timeout)));
	}
	conn->txused = (int) sizeof(CbIcPktHdr);
}

/* before this process blocks in a receive: everything it has buffered for others must be on its way */
static void
cbic_flush_all(ChunkTransportState *ts, CbIcState *st)
{
	for (int i = 0; i < st->nnodes; i++)
		if (st->nodes[i].is_sender)
			for (int c = 0; c < st->nodes[i].nconns; c++)
				if (st->nodes[i].conns[c].still_active)
					cbic_flush(ts, st, &st->nodes[i], &st->nodes[i].conns[c], 0);
}

static void
cbic_append_chunk(ChunkTransportState *ts, CbIcState *st, CbIcNode *node, CbIcConn *conn, TupleChunkListItem item)
{
	if ((int) item->chunk_length > st->max_packet - (int) sizeof(CbIcPktHdr))
		ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: tuple chunk of %u bytes exceeds the packet size", item->chunk_length)));
	if (conn->txused + (int) item->chunk_length > st->max_packet)
		cbic_flush(ts, st, node, conn, 0);
	memcpy(conn->tx + conn->txused, GetChunkDataPtr(item), item->chunk_length);
	conn->txused += (int) item->chunk_length;
}

/* the chunks of one received packet as a palloc'ed TupleChunkListItem chain (ml_ipc.h:170: "allocated with palloc()").  The
 * items point INTO the packet (`inplace`), as the socket interconnects' point into their receive buffers: cdbmotion.c copies
 * only the chunks it must keep (materializeChunk, :953) and gives the buffer back with DirectPutRxBuffer (:700). */
static TupleChunkListItem
cbic_packet_to_chunks(CbIcPkt *p)
{
	TupleChunkListItem first = NULL,
				last = NULL;
	uint32		pos = 0;

	while (pos + TUPLE_CHUNK_HEADER_SIZE <= p->nbytes)
	{
		uint16		size;
		uint32		len;
		TupleChunkListItem it;

		memcpy(&size, p->data + pos, sizeof(uint16));	/* GetChunkDataSize: the first header field */
		len = TUPLE_CHUNK_HEADER_SIZE + size;
		if (pos + len > p->nbytes)
			ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: chunk runs past its packet")));
		it = palloc(sizeof(TupleChunkListItemData));
		it->p_next = NULL;
		it->chunk_length = len;
		it->inplace = (char *) p->data + pos;
		if (last)
			last->p_next = it;
		else
			first = it;
		last = it;
		pos += len;
	}
	return first;
}

/* ------------------------------------------------------------------------------------------
 * the vtable
 * ------------------------------------------------------------------------------------------ */
static int
cbic_GetMaxTupleChunkSize(void)
{
	return Gp_max_packet_size - PACKET_HEADER_SIZE - (int) sizeof(CbIcPktHdr);
}

static int32
cbic_GetListenPort(void)
{
	return 0;					/* no listener: endpoints find each other by name (file header) */
}

static void
cbic_InitMotionLayerIPC(void)
{
}

static void
cbic_CleanUpMotionLayerIPC(void)
{
}

static void
cbic_WaitInterconnectQuit(void)
{
}

static int
cbic_slice_base(SliceTable *tab, int slice)
{
	int			base = 0;

	for (int k = 0; k < slice; k++)
		base += list_length(tab->slices[k].primaryProcesses);
	return base;
}

static void
cbic_init_conns(CbIcState *st, CbIcNode *node, SliceTable *tab, ExecSlice *peer, bool sender)
{
	node->is_sender = sender;
	node->nconns = list_length(peer->primaryProcesses);
	node->conns = MemoryContextAllocZero(st->cxt, sizeof(CbIcConn) * Max(node->nconns, 1));
	for (int i = 0; i < node->nconns; i++)
	{
		CbIcConn   *c = &node->conns[i];

		c->rank = cbic_slice_base(tab, peer->sliceIndex) + i;
		c->still_active = sender && list_nth(peer->primaryProcesses, i) != NULL;
		if (sender)
		{
			c->tx = MemoryContextAlloc(st->cxt, st->max_packet);
			c->txused = (int) sizeof(CbIcPktHdr);
		}
		else
			c->remapper = CreateTupleRemapper();
		cbic_active_conns++;
	}
}

static void
cbic_SetupInterconnect(struct EState *estate)
{
	SliceTable *tab = estate->es_sliceTable;
	ExecSlice  *mine;
	ChunkTransportState *ts;
	CbIcState  *st;
	MemoryContext old;
	ListCell   *lc;
	char		token[96];
	CbChanMem	mem;
	int64		waited = 0;

	if (estate->interconnect_context)
		ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: already set up for this statement")));
	if (tab == NULL || tab->localSlice < 0 || tab->localSlice >= tab->numSlices)
		ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: no slice table")));
	mine = &tab->slices[tab->localSlice];
	old = MemoryContextSwitchTo(estate->es_query_cxt);
	ts = palloc0(sizeof(ChunkTransportState));
	st = palloc0(sizeof(CbIcState));
	st->cxt = estate->es_query_cxt;
	ts->estate = estate;
	ts->sliceTable = tab;
	ts->sliceId = tab->localSlice;
	ts->implement_state = st;
	/* endpoint numbering: every process of every slice, in slice-table order */
	st->nranks = cbic_slice_base(tab, tab->numSlices);
	st->my_route = -1;
	{
		int			i = 0;

		foreach(lc, mine->primaryProcesses)
		{
			CdbProcess *p = (CdbProcess *) lfirst(lc);

			if (p != NULL && p->pid == MyProcPid)
				st->my_route = i;
			i++;
		}
	}
	if (st->my_route < 0)
		ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: this process (pid %d) is not in slice %d", MyProcPid, mine->sliceIndex)));
	st->rank = cbic_slice_base(tab, mine->sliceIndex) + st->my_route;
	st->max_packet = CBIC_SLOT_BYTES - 8;
	if (Gp_max_packet_size > st->max_packet)
		ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: gp_max_packet_size %d exceeds the channel's packets (%d)", Gp_max_packet_size, st->max_packet)));
	st->rxbuf = palloc(st->max_packet);

	/* my arena, then everybody else's (waiting for late starters, cancellably) */
	snprintf(token, sizeof(token), "cbic_%d_%d", gp_session_id, gp_command_count);
	st->shm = cb_chan_shm_create(token, st->rank, st->nranks, cb_chan_arena_bytes(st->nranks, CBIC_SLOTS, CBIC_SLOT_BYTES));
	if (st->shm == NULL)
		ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: cannot create the arena of endpoint %d: %m", st->rank)));
	while (cb_chan_shm_attach(st->shm, &mem) != 0)
	{
		CHECK_FOR_INTERRUPTS();
		pg_usleep(2000);
		waited += 2;
		if (interconnect_setup_timeout > 0 && waited > (int64) interconnect_setup_timeout * 1000)
		{
			cb_chan_shm_close(st->shm, 1);
			ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: not every endpoint appeared within %d s", interconnect_setup_timeout)));
		}
	}
	st->chan = cb_chan_create(st->rank, st->nranks, CBIC_SLOTS, CBIC_SLOT_BYTES, &mem);
	if (st->chan == NULL)
		ereport(ERROR, (errcode(ERRCODE_OUT_OF_MEMORY), errmsg("cbgpu interconnect: out of memory")));

	/* the Motions this process takes part in: receiver of every child slice's, sender of its own slice's */
	st->nodes = palloc0(sizeof(CbIcNode) * (list_length(mine->children) + 1));
	foreach(lc, mine->children)
	{
		ExecSlice  *child = &tab->slices[lfirst_int(lc)];
		CbIcNode   *node = &st->nodes[st->nnodes++];

		node->motNodeID = (int16) child->sliceIndex;
		cbic_init_conns(st, node, tab, child, false);
	}
	if (mine->parentIndex >= 0)
	{
		CbIcNode   *node = &st->nodes[st->nnodes++];

		node->motNodeID = (int16) mine->sliceIndex;
		cbic_init_conns(st, node, tab, &tab->slices[mine->parentIndex], true);
	}
	ts->activated = true;
	estate->interconnect_context = ts;
	estate->es_interconnect_is_setup = true;
	MemoryContextSwitchTo(old);
}

static void
cbic_TeardownInterconnect(ChunkTransportState *ts, bool hasErrors)
{
	CbIcState  *st;

	if (ts == NULL || ts->implement_state == NULL)
		return;
	st = (CbIcState *) ts->implement_state;
	ts->teardownActive = true;
	if (!hasErrors)
	{
		/* whatever is still buffered goes out; a sender that never reached SendEOS (squelched) is the executor's business */
		PG_TRY();
		{
			cbic_flush_all(ts, st);
		}
		PG_CATCH();
		{
			FlushErrorState();
		}
		PG_END_TRY();
	}
	for (int i = 0; i < st->nnodes; i++)
	{
		cbic_active_conns -= st->nodes[i].nconns;
		for (int c = 0; c < st->nodes[i].nconns; c++)
			if (st->nodes[i].conns[c].remapper)
				DestroyTupleRemapper(st->nodes[i].conns[c].remapper);
	}
	if (st->chan)
		cb_chan_destroy(st->chan);
	if (st->shm)
		cb_chan_shm_close(st->shm, 1);	/* peers that still have it mapped keep their mapping; the name goes now */
	st->chan = NULL;
	st->shm = NULL;
	ts->implement_state = NULL;
	ts->activated = false;
}

static bool
cbic_SendTupleChunkToAMS(ChunkTransportState *ts, int16 motNodeID, int16 targetRoute, TupleChunkListItem tcItem)
{
	CbIcState  *st = cbic_state(ts);
	CbIcNode   *node = cbic_node(st, motNodeID, true);
	bool		any_active = false;

	ML_CHECK_FOR_INTERRUPTS(ts->teardownActive);
	if (targetRoute != BROADCAST_SEGIDX && (targetRoute < 0 || targetRoute >= node->nconns))
		ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: targetRoute %d outside 0 .. %d", targetRoute, node->nconns - 1)));
	for (TupleChunkListItem it = tcItem; it != NULL; it = it->p_next)
		for (int c = 0; c < node->nconns; c++)
			if ((targetRoute == BROADCAST_SEGIDX || c == targetRoute) && node->conns[c].still_active)
				cbic_append_chunk(ts, st, node, &node->conns[c], it);
	/* stop messages that arrived meanwhile take effect here (the reference polls them in SendChunk) */
	if (cb_chan_pending(st->chan, -1) > 0)
		cbic_drain(st, 0);
	for (int c = 0; c < node->nconns; c++)
		any_active |= node->conns[c].still_active;
	return any_active;			/* false: every receiver has stopped us (cdbmotion.c:405-408 -> STOP_SENDING) */
}

static bool
cbic_SendChunk(ChunkTransportState *ts, ChunkTransportStateEntry *pEntry, MotionConn *conn, TupleChunkListItem tcItem, int16 motionId)
{
	(void) ts; (void) pEntry; (void) conn; (void) tcItem; (void) motionId;
	ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: SendChunk is internal to the socket interconnects")));
	return false;
}

static void
cbic_SendEOS(ChunkTransportState *ts, int motNodeID, TupleChunkListItem tcItem)
{
	CbIcState  *st = cbic_state(ts);
	CbIcNode   *node = cbic_node(st, motNodeID, true);

	for (int c = 0; c < node->nconns; c++)
	{
		CbIcConn   *conn = &node->conns[c];

		if (conn->eos_sent)
			continue;
		if (conn->still_active)
			cbic_append_chunk(ts, st, node, conn, tcItem);
		else
			conn->txused = (int) sizeof(CbIcPktHdr);	/* a stopped receiver still learns that the stream ended */
		cbic_flush(ts, st, node, conn, CBIC_FLAG_EOS);
		conn->eos_sent = true;
		conn->still_active = false;
	}
}

static void
cbic_SendStopMessage(ChunkTransportState *ts, int16 motNodeID)
{
	CbIcState  *st = cbic_state(ts);
	CbIcNode   *node = cbic_node(st, motNodeID, false);
	CbIcPktHdr	h;

	h.magic = CBIC_MAGIC;
	h.motNodeID = motNodeID;
	h.srcRoute = (int16) st->my_route;
	h.flags = CBIC_FLAG_STOP;
	h.nbytes = 0;
	h.seq = 0;
	for (int c = 0; c < node->nconns; c++)
		if (!node->conns[c].eos_seen)
			(void) cb_chan_send(st->chan, node->conns[c].rank, &h, (int) sizeof(h), 1000);	/* best effort, as UDP's is */
}

static TupleChunkListItem
cbic_recv(ChunkTransportState *ts, int16 motNodeID, int16 *srcRoute, bool any)
{
	CbIcState  *st = cbic_state(ts);
	CbIcNode   *node = cbic_node(st, motNodeID, false);
	int64		waited = 0;

	for (;;)
	{
		bool		open_stream = false;

		/* a queued packet?  Any-source receives scan the queues from where the last one ended (fairness, ml_ipc.h:170-178) */
		for (int k = 0; k < node->nconns; k++)
		{
			const int	c = any ? (node->next_any + k) % node->nconns : *srcRoute;
			CbIcConn   *conn = &node->conns[c];

			if (!any && k > 0)
				break;
			if (conn->deregistered)
				continue;
			if (conn->rx_head)
			{
				CbIcPkt    *p = conn->rx_head;
				TupleChunkListItem items;

				conn->rx_head = p->next;
				if (conn->rx_head == NULL)
					conn->rx_tail = NULL;
				if (p->flags & CBIC_FLAG_EOS)
					conn->eos_seen = true;
				if (conn->rx_held)
					pfree(conn->rx_held);	/* a caller that skipped DirectPutRxBuffer is done with it by now */
				conn->rx_held = p;
				items = cbic_packet_to_chunks(p);
				if (any)
				{
					node->next_any = (c + 1) % node->nconns;
					*srcRoute = (int16) c;
				}
				if (items)
					return items;
				continue;		/* an EOS packet of a stopped stream may carry nothing */
			}
			if (!conn->eos_seen)
				open_stream = true;
		}
		if (!open_stream)
			ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR),
							errmsg("cbgpu interconnect: receive on motion node %d after every sender has ended its stream", motNodeID)));
		/* nothing queued: about to wait.  My own buffered sends first - a peer may be waiting for exactly those */
		cbic_flush_all(ts, st);
		ML_CHECK_FOR_INTERRUPTS(ts->teardownActive);
		if (cbic_drain(st, 50) == 0)
		{
			waited += 50;
			if (Gp_interconnect_transmit_timeout > 0 && waited > (int64) Gp_interconnect_transmit_timeout * 1000)
				ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR),
								errmsg("cbgpu interconnect: no packet for motion node %d in %d s", motNodeID, Gp_interconnect_transmit_timeout)));
		}
		else
			waited = 0;
	}
}

static TupleChunkListItem
cbic_RecvTupleChunkFromAny(ChunkTransportState *ts, int16 motNodeID, int16 *srcRoute)
{
	return cbic_recv(ts, motNodeID, srcRoute, true);
}

static TupleChunkListItem
cbic_RecvTupleChunkFrom(ChunkTransportState *ts, int16 motNodeID, int16 srcRoute)
{
	CbIcState  *st = cbic_state(ts);
	CbIcNode   *node = cbic_node(st, motNodeID, false);

	if (srcRoute < 0 || srcRoute >= node->nconns)
		ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: srcRoute %d outside 0 .. %d", srcRoute, node->nconns - 1)));
	return cbic_recv(ts, motNodeID, &srcRoute, false);
}

static TupleChunkListItem
cbic_RecvTupleChunk(MotionConn *conn, ChunkTransportState *ts)
{
	(void) conn; (void) ts;
	ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: RecvTupleChunk is internal to the socket interconnects")));
	return NULL;
}

static void
cbic_DirectPutRxBuffer(ChunkTransportState *ts, int motNodeID, int route)
{
	/* the caller has processed the chunk list of its last receive on this route: the packet they pointed into can go */
	CbIcNode   *node = cbic_node(cbic_state(ts), motNodeID, false);

	if (route >= 0 && route < node->nconns && node->conns[route].rx_held)
	{
		pfree(node->conns[route].rx_held);
		node->conns[route].rx_held = NULL;
	}
}

static void
cbic_DeregisterReadInterest(ChunkTransportState *ts, int motNodeID, int srcRoute, const char *reason)
{
	CbIcState  *st = cbic_state(ts);
	CbIcNode   *node = cbic_node(st, motNodeID, false);

	(void) reason;
	if (srcRoute < 0 || srcRoute >= node->nconns)
		return;
	node->conns[srcRoute].deregistered = true;
	while (node->conns[srcRoute].rx_head)
	{
		CbIcPkt    *p = node->conns[srcRoute].rx_head;

		node->conns[srcRoute].rx_head = p->next;
		pfree(p);
	}
	node->conns[srcRoute].rx_tail = NULL;
}

static uint32
cbic_GetActiveMotionConns(void)
{
	return (uint32) Max(cbic_active_conns, 0);
}

static void
cbic_GetTransportDirectBuffer(ChunkTransportState *ts, int16 motNodeID, int16 targetRoute, struct directTransportBuffer *b)
{
	/* tuples go through the chunk list (SerializeTuple falls back to it when no direct buffer is offered, tupser.c:349):
	 * the transmit buffer belongs to the packet being assembled, whose header is written at flush time */
	(void) ts; (void) motNodeID; (void) targetRoute;
	b->pri = NULL;
	b->prilen = 0;
}

static void
cbic_PutTransportDirectBuffer(ChunkTransportState *ts, int16 motNodeID, int16 targetRoute, int serializedLength)
{
	(void) ts; (void) motNodeID; (void) targetRoute;
	if (serializedLength != 0)
		ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: no direct transmit buffer was handed out")));
}

static TupleRemapper *
cbic_GetMotionConnTupleRemapper(ChunkTransportState *ts, int16 motNodeID, int16 targetRoute)
{
	CbIcNode   *node = cbic_node(cbic_state(ts), motNodeID, false);

	if (targetRoute < 0 || targetRoute >= node->nconns)
		ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: route %d outside 0 .. %d", targetRoute, node->nconns - 1)));
	return node->conns[targetRoute].remapper;
}

static int32 *
cbic_GetMotionSentRecordTypmod(ChunkTransportState *ts, int16 motNodeID, int16 targetRoute)
{
	CbIcNode   *node = cbic_node(cbic_state(ts), motNodeID, true);

	if (targetRoute == BROADCAST_SEGIDX)
		targetRoute = 0;		/* the reference keeps ONE counter for a broadcast: conns[0]'s (ic_common.c:540-555) */
	if (targetRoute < 0 || targetRoute >= node->nconns)
		ereport(ERROR, (errcode(ERRCODE_GP_INTERCONNECTION_ERROR), errmsg("cbgpu interconnect: route %d outside 0 .. %d", targetRoute, node->nconns - 1)));
	return &node->conns[targetRoute].sent_record_typmod;
}

MotionIPCLayer cbgpu_ipc_layer = {
	.ic_type = INTERCONNECT_TYPE_UDP2,	/* the slot of a module this cluster does not load (file header) */
	.type_name = "cbgpu",
	.GetMaxTupleChunkSize = cbic_GetMaxTupleChunkSize,
	.GetListenPort = cbic_GetListenPort,
	.InitMotionLayerIPC = cbic_InitMotionLayerIPC,
	.CleanUpMotionLayerIPC = cbic_CleanUpMotionLayerIPC,
	.WaitInterconnectQuit = cbic_WaitInterconnectQuit,
	.SetupInterconnect = cbic_SetupInterconnect,
	.TeardownInterconnect = cbic_TeardownInterconnect,
	.SendTupleChunkToAMS = cbic_SendTupleChunkToAMS,
	.SendChunk = cbic_SendChunk,
	.SendEOS = cbic_SendEOS,
	.SendStopMessage = cbic_SendStopMessage,
	.RecvTupleChunkFromAny = cbic_RecvTupleChunkFromAny,
	.RecvTupleChunkFrom = cbic_RecvTupleChunkFrom,
	.RecvTupleChunk = cbic_RecvTupleChunk,
	.DirectPutRxBuffer = cbic_DirectPutRxBuffer,
	.DeregisterReadInterest = cbic_DeregisterReadInterest,
	.GetActiveMotionConns = cbic_GetActiveMotionConns,
	.GetTransportDirectBuffer = cbic_GetTransportDirectBuffer,
	.PutTransportDirectBuffer = cbic_PutTransportDirectBuffer,
	.IcProxyServiceMain = NULL,
	.GetMotionConnTupleRemapper = cbic_GetMotionConnTupleRemapper,
	.GetMotionSentRecordTypmod = cbic_GetMotionSentRecordTypmod,
};

#ifndef CBGPU_IC_NO_MODULE_MAGIC
void		_PG_init(void);

void
_PG_init(void)
{
	if (!process_shared_preload_libraries_in_progress)
		ereport(ERROR, (errcode_for_file_access(), errmsg("could not load the cbgpu interconnect outside process shared preload")));
	RegisterIPCLayerImpl(&cbgpu_ipc_layer);
}
#endif
