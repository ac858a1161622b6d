/*
 * cbgpu_shim_motion.c - GPU rows straight onto the backend's own interconnect, and off it.
 *
 * When the GPU sub-tree ends below a Motion node that stays with the reference (its receivers are CPU slices), the
 * plain route already works: the Motion pulls virtual slots from the replaced ExecProcNode and SendTuple() serialises
 * them (cdb/motion/cdbmotion.c:1315-1400).  This file is the bulk route around the slots: rows the GPU executor
 * hands over as values go through cb_tupser_row (include/cb_exec.h; byte-identical to SerializeTuple, see
 * tests/test_tupser.py) and the resulting chunks through the current MotionIPCLayer's SendTupleChunkToAMS
 * (include/cdb/ml_ipc.h:123) -- whichever interconnect is configured (udpifc, tcp, proxy).  The other direction reads
 * chunks with RecvTupleChunkFromAny (:188) and gives executor values back (cb_tupser_next), ready for
 * cbgpu_rel_load_column.
 *
 * Type-checked against the reference's headers (tests/test_shim_compiles.py); not linked or run here (no backend).
 */
#include "postgres.h"

#include "cdb/cdbmotion.h"
#include "cdb/ml_ipc.h"
#include "cdb/tupchunklist.h"
#include "miscadmin.h"
#include "nodes/execnodes.h"
#include "nodes/plannodes.h"

#include "cb_exec.h"

/*
 * Send nrows rows (row-major values / isnull, as cb_tupser_row takes them) of Motion `motion`; route[r] is the receiver
 * of row r (BROADCAST_SEGIDX for a Broadcast Motion).  false = a receiver asked us to stop (SendStopMessage), as
 * SendTuple reports with STOP_SENDING.
 */
bool
cbgpu_shim_send_rows(EState *estate, Motion *motion, const CbTupAttr *attrs, int natts, const int64 *values, const uint8 *isnull,
					 int64 nrows, const int16 *route)
{
	ChunkTransportState *ts = estate->interconnect_context;
	const int	cap = Gp_max_tuple_chunk_size;
	/* room for the longest row this executor produces: a few chunks; grown on demand */
	int64		bufcap = 4 * (int64) cap;
	unsigned char *buf = palloc(bufcap);
	TupleChunkListItem item = palloc0(sizeof(TupleChunkListItemData));

	for (int64 r = 0; r < nrows; r++)
	{
		int64		n;
		int64		pos = 0;

		CHECK_FOR_INTERRUPTS();
		while ((n = cb_tupser_row(attrs, natts, values + r * natts, isnull ? isnull + r * natts : NULL, cap, buf, bufcap)) == -2)
		{
			bufcap *= 2;
			buf = repalloc(buf, bufcap);
		}
		if (n < 0)
			ereport(ERROR, (errcode(ERRCODE_INTERNAL_ERROR), errmsg("cbgpu: row " INT64_FORMAT " cannot be serialised", r)));
		while (pos < n)
		{
			uint16		size;

			memcpy(&size, buf + pos, sizeof(uint16));
			/* a chunk "in place": the transport copies TUPLE_CHUNK_HEADER_SIZE + size bytes from it (GetChunkDataPtr) */
			item->p_next = NULL;
			item->chunk_length = TUPLE_CHUNK_HEADER_SIZE + size;
			item->inplace = (char *) buf + pos;
			if (!CurrentMotionIPCLayer->SendTupleChunkToAMS(ts, motion->motionID, route[r], item))
			{
				pfree(buf);
				pfree(item);
				return false;
			}
			pos += TUPLE_CHUNK_HEADER_SIZE + size;
		}
	}
	pfree(buf);
	pfree(item);
	return true;
}

/* the sender is done: SendEndOfStream (cdb/motion/cdbmotion.c:1160) builds and broadcasts the TC_END_OF_STREAM chunk */
void
cbgpu_shim_send_end(EState *estate, Motion *motion)
{
	SendEndOfStream((MotionLayerState *) estate->motionlayer_context, estate->interconnect_context, motion->motionID);
}

/*
 * Receive up to maxrows rows of Motion `motion` from any sender into values / isnull (row-major).  Returns the rows
 * read; *eos is set once every sender has closed its stream (nsenders TC_END_OF_STREAM chunks seen, counted in *ended,
 * which the caller keeps across calls, starting at 0).
 */
int64
cbgpu_shim_recv_rows(EState *estate, Motion *motion, int nsenders, const CbTupAttr *attrs, int natts, int64 *values, uint8 *isnull,
					 int64 maxrows, int *ended, bool *eos)
{
	ChunkTransportState *ts = estate->interconnect_context;
	StringInfoData pending;
	int64		rows = 0;

	/* chunks of one tuple arrive back to back from one sender (tuple chunks carry no tuple id: cdb/tupchunklist.h), so a
	 * per-route reassembly buffer is what a full implementation keeps; with a single pending buffer this sketch handles
	 * the common case of rows that fit one chunk plus in-order partial chunks from one route at a time */
	initStringInfo(&pending);
	*eos = false;
	/* leave room for what one read can deliver: a packet of Gp_max_packet_size holds at most a few hundred tuples */
	while (rows + 1024 < maxrows && *ended < nsenders)
	{
		int16		src = ANY_ROUTE;
		TupleChunkListItem it = CurrentMotionIPCLayer->RecvTupleChunkFromAny(ts, motion->motionID, &src);

		for (; it != NULL; it = it->p_next)
		{
			int64		used = 0;
			int64		rc;

			appendBinaryStringInfo(&pending, GetChunkDataPtr(it), (int) it->chunk_length);
			if (rows >= maxrows)
				ereport(ERROR, (errcode(ERRCODE_INTERNAL_ERROR), errmsg("cbgpu: receive batch too small")));
			rc = cb_tupser_next(attrs, natts, (const unsigned char *) pending.data, pending.len, &used, values + rows * natts,
								isnull + rows * natts);
			if (rc == CB_TUPSER_NEED_MORE)
				continue;
			if (rc == CB_TUPSER_BAD)
				ereport(ERROR, (errcode(ERRCODE_PROTOCOL_VIOLATION), errmsg("cbgpu: malformed tuple chunk from route %d", src)));
			if (rc == CB_TUPSER_END)
				(*ended)++;
			else if (++rows > maxrows)
				ereport(ERROR, (errcode(ERRCODE_INTERNAL_ERROR),
								errmsg("cbgpu: one interconnect read delivered more rows than the batch holds (" INT64_FORMAT ")", maxrows)));
			resetStringInfo(&pending);
		}
	}
	*eos = *ended >= nsenders;
	pfree(pending.data);
	return rows;
}
