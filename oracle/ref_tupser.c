/*
 * ref_tupser.c - drives the REFERENCE's own Motion tuple serialisation for the oracle (test infrastructure):
 * heap_form_minimal_tuple / heap_deform_tuple (access/common/heaptuple.c), SerializeTuple / CvtChunksToTup
 * (cdb/motion/tupser.c) and the chunk list (cdb/motion/tupchunklist.c), compiled where they lie into
 * oracle/_ref/libaocs_ref.so.  What a Motion sender puts on the wire for a row, and what a receiver makes of chunks.
 *
 * Backend pieces those files call but this path never reaches are stubs that abort; palloc / ereport come from
 * ref_aocs.c.  A virtual TupleTableSlot is assembled by hand (execTuples.c is not compiled): fetching its minimal
 * tuple is heap_form_minimal_tuple over tts_values / tts_isnull, which is what tts_virtual_copy_minimal_tuple does
 * (executor/execTuples.c).
 */
#include "postgres.h"

#include "access/htup_details.h"
#include "access/tupdesc.h"
#include "cdb/tupchunklist.h"
#include "cdb/tupser.h"
#include "cdb/tupleremap.h"
#include "executor/tuptable.h"

int			Gp_max_tuple_chunk_size = 8192 - 32;
const TupleTableSlotOps TTSOpsVirtual;

MinimalTuple
ExecFetchSlotMinimalTuple(TupleTableSlot *slot, bool *shouldFree)
{
	*shouldFree = true;
	return heap_form_minimal_tuple(slot->tts_tupleDescriptor, slot->tts_values, slot->tts_isnull);
}

static TupleDesc
make_desc(int natts, const int *typid, const int *typlen, const int *byval, const int *align, const int *storage)
{
	TupleDesc	d = (TupleDesc) palloc0(offsetof(struct TupleDescData, attrs) + natts * sizeof(FormData_pg_attribute));

	d->natts = natts;
	d->tdrefcount = -1;
	for (int i = 0; i < natts; i++)
	{
		Form_pg_attribute a = TupleDescAttr(d, i);

		a->atttypid = (Oid) typid[i];
		a->attlen = (int16) typlen[i];
		a->attbyval = byval[i] != 0;
		a->attalign = (char) align[i];
		a->attstorage = (char) storage[i];
		a->attnum = (int16) (i + 1);
		a->atttypmod = -1;
		a->attcacheoff = -1;
	}
	return d;
}

/* the jmp_buf ereport(ERROR) unwinds to, and the error text, live in ref_aocs.c */
#include <setjmp.h>
extern jmp_buf ref_jmp;

/*
 * Serialise nrows rows (values row-major: by-value datums, or offsets into varbuf of 4-byte-header varlenas) the way a
 * Motion sender does with the direct transport buffer unavailable: every tuple's chunks back to back.  Returns bytes.
 */
int64
ref_tupser_serialize(int natts, const int *typid, const int *typlen, const int *byval, const int *align, const int *storage,
					 const int64 *values, const unsigned char *varbuf, const unsigned char *nulls, int64 nrows, int max_chunk,
					 unsigned char *out, int64 outcap, int64 *nchunks_out)
{
	TupleDesc	desc;
	SerTupInfo	ser;
	TupleTableSlot slot;
	struct directTransportBuffer { char *pri; int prilen; } b = {NULL, 0};
	Datum	   *vals;
	bool	   *isnull;
	int64		pos = 0,
				nchunks = 0;

	if (setjmp(ref_jmp))
		return -1;
	Gp_max_tuple_chunk_size = max_chunk;
	desc = make_desc(natts, typid, typlen, byval, align, storage);
	memset(&ser, 0, sizeof(ser));
	ser.tupdesc = desc;
	vals = palloc(sizeof(Datum) * (natts ? natts : 1));
	isnull = palloc(sizeof(bool) * (natts ? natts : 1));
	memset(&slot, 0, sizeof(slot));
	slot.type = T_TupleTableSlot;
	*(const TupleTableSlotOps **) &slot.tts_ops = &TTSOpsVirtual;	/* the member is const: set once, as MakeTupleTableSlot does */
	slot.tts_tupleDescriptor = desc;
	slot.tts_values = vals;
	slot.tts_isnull = isnull;
	slot.tts_nvalid = (AttrNumber) natts;
	for (int64 r = 0; r < nrows; r++)
	{
		TupleChunkListData tc;
		TupleChunkListItem it;

		for (int a = 0; a < natts; a++)
		{
			isnull[a] = nulls && nulls[r * natts + a];
			vals[a] = isnull[a] ? (Datum) 0 : typlen[a] == -1 ? PointerGetDatum(varbuf + values[r * natts + a]) : (Datum) values[r * natts + a];
		}
		memset(&tc, 0, sizeof(tc));
		if (SerializeTuple(&slot, &ser, (struct directTransportBuffer *) &b, &tc, 0) != 0)
			return -2;
		for (it = tc.p_first; it; it = it->p_next)
		{
			if (pos + it->chunk_length > outcap)
				return -3;
			memcpy(out + pos, it->chunk_data, it->chunk_length);
			pos += it->chunk_length;
			nchunks++;
		}
		clearTCList(NULL, &tc);
	}
	if (nchunks_out)
		*nchunks_out = nchunks;
	return pos;
}

/*
 * The receiving side over a byte stream of chunks: CvtChunksToTup per tuple, heap_deform_tuple.  values_out row-major
 * (varlena attributes: offset into varbuf_out, where the datum is copied with whatever header it arrived with);
 * returns rows, -1 on error.
 */
int64
ref_tupser_deserialize(int natts, const int *typid, const int *typlen, const int *byval, const int *align, const int *storage,
					   const unsigned char *chunks, int64 nbytes, int64 *values_out, unsigned char *nulls_out, int64 maxrows,
					   unsigned char *varbuf_out, int64 varcap)
{
	TupleDesc	desc;
	SerTupInfo	ser;
	int64		pos = 0,
				rows = 0,
				vpos = 0;
	Datum	   *vals;
	bool	   *isnull;

	if (setjmp(ref_jmp))
		return -1;
	desc = make_desc(natts, typid, typlen, byval, align, storage);
	memset(&ser, 0, sizeof(ser));
	ser.tupdesc = desc;
	vals = palloc(sizeof(Datum) * (natts ? natts : 1));
	isnull = palloc(sizeof(bool) * (natts ? natts : 1));
	while (pos < nbytes)
	{
		TupleChunkListData tc;
		MinimalTuple mt;
		HeapTupleData htup;
		bool		done = false;

		memset(&tc, 0, sizeof(tc));
		while (!done)
		{
			uint16		size,
						type;
			TupleChunkListItem it;

			if (pos + TUPLE_CHUNK_HEADER_SIZE > nbytes)
				return -2;
			memcpy(&size, chunks + pos, 2);
			memcpy(&type, chunks + pos + 2, 2);
			if (pos + TUPLE_CHUNK_HEADER_SIZE + size > nbytes)
				return -2;
			if (type == TC_END_OF_STREAM)
				return rows;
			it = palloc0(sizeof(TupleChunkListItemData) + TUPLE_CHUNK_HEADER_SIZE);
			it->chunk_length = TUPLE_CHUNK_HEADER_SIZE + size;
			it->inplace = (char *) chunks + pos;
			appendChunkToTCList(&tc, it);
			pos += TUPLE_CHUNK_HEADER_SIZE + size;
			done = type == TC_WHOLE || type == TC_PARTIAL_END || type == TC_EMPTY;
		}
		mt = CvtChunksToTup(&tc, &ser, NULL);
		if (mt == NULL || rows >= maxrows)
			return -3;
		htup.t_len = mt->t_len + MINIMAL_TUPLE_OFFSET;
		htup.t_data = (HeapTupleHeader) ((char *) mt - MINIMAL_TUPLE_OFFSET);
		heap_deform_tuple(&htup, desc, vals, isnull);
		for (int a = 0; a < natts; a++)
		{
			nulls_out[rows * natts + a] = isnull[a] ? 1 : 0;
			if (isnull[a])
				values_out[rows * natts + a] = 0;
			else if (typlen[a] == -1)
			{
				const Size	sz = VARSIZE_ANY(DatumGetPointer(vals[a]));

				if (vpos + (int64) sz > varcap)
					return -4;
				memcpy(varbuf_out + vpos, DatumGetPointer(vals[a]), sz);
				values_out[rows * natts + a] = vpos;
				vpos += (int64) sz;
			}
			else
				values_out[rows * natts + a] = (int64) vals[a];
		}
		rows++;
	}
	return rows;
}
