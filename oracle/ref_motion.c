/*
 * ref_motion.c - drives the REFERENCE's Motion layer (cdb/motion/cdbmotion.c: SendTuple / SendEndOfStream / RecvTupleFrom,
 * with tupser.c, tupchunklist.c, htupfifo.c, access/common/heaptuple.c, nodes/list.c compiled where they lie) over the
 * product's MotionIPCLayer implementation integration/cbgpu_ic_layer.c between real processes.  Test infrastructure
 * (oracle/_ref/libmotion_ref.so, tests/test_ic_layer.py): one call = one QE process of a two-slice plan
 *
 *        slice 0 (receivers, root)   <-   Motion 1   <-   slice 1 (senders)
 *
 * that sets the interconnect up through the layer's vtable exactly as the executor does (SetupInterconnect with the slice
 * table, execMain.c:531), then sends its rows with SendTuple (per-row target route, or BROADCAST_SEGIDX) and closes with
 * SendEndOfStream, or receives with RecvTupleFrom(ANY_ROUTE) until every sender's stream has ended.
 *
 * Backend pieces this path reaches are restated minimally (memory contexts over malloc, a TupleDesc copy, the GUC
 * variables); pieces it never reaches abort (ref_stubs.c).  palloc / ereport come from ref_aocs.c.
 */
#include "postgres.h"

#include <setjmp.h>

#include "access/htup_details.h"
#include "access/tupdesc.h"
#include "cdb/cdbgang.h"
#include "cdb/cdbinterconnect.h"
#include "cdb/cdbmotion.h"
#include "cdb/cdbvars.h"
#include "cdb/ml_ipc.h"
#include "access/session.h"
#include "catalog/pg_type.h"
#include "cdb/tupleremap.h"
#include "executor/tuptable.h"
#include "nodes/execnodes.h"
#include "nodes/pg_list.h"
#include "utils/memutils.h"

extern jmp_buf ref_jmp;
extern MotionIPCLayer cbgpu_ipc_layer;
extern const TupleTableSlotOps TTSOpsVirtual;

/* ---- what cdbmotion.c and the layer read from the backend ---- */
GpRoleValue Gp_role = GP_ROLE_EXECUTE;
int			Gp_max_packet_size = 8192;
int			Gp_interconnect_transmit_timeout = 60;
int			interconnect_setup_timeout = 60;
int			gp_session_id = 1;
int			gp_command_count = 1;
int			gp_log_interconnect = 0;
int			gp_motion_slice_noop = 0;
int			currentSliceId = 0;
int			MyProcPid = 0;
bool		process_shared_preload_libraries_done = true;
volatile sig_atomic_t InterruptPending = false;
int32		NextRecordTypmod = 0;
static Session the_session;	/* no shared record-typmod registry: NextRecordTypmod (0) rules, nothing is sent */
Session    *CurrentSession = &the_session;
MemoryContext CurrentMemoryContext = NULL;
MemoryContext TopMemoryContext = NULL;
sigjmp_buf *PG_exception_stack = NULL;
ErrorContextCallback *error_context_stack = NULL;

int			Gp_interconnect_type = 0;
/* CHECK_FOR_INTERRUPTS' riders (miscadmin.h INTERRUPTS_MORE_CHECK): resource-queue backoff, OOM report, runaway detection */
int			backoffTickCounter = 0;
int			gp_resqueue_priority_local_interval = 1 << 30;
int			dispatch_nest_level = 0;
bool		gp_mp_inited = false;
volatile OOMTimeType *segmentOOMTime = NULL;
volatile OOMTimeType oomTrackerStartTime = 0;
volatile OOMTimeType alreadyReportedOOMTime = 0;
void		BackoffBackendTickExpired(void) { backoffTickCounter = 0; }
void		RedZoneHandler_DetectRunawaySession(void) {}
void		UpdateTimeAtomically(volatile OOMTimeType *t) { (void) t; }
void		MemoryContextStats(MemoryContext c) { (void) c; }
void		write_stderr(const char *fmt,...) { (void) fmt; }
void		ProcessInterrupts(const char *filename, int lineno) { (void) filename; (void) lineno; }
void		pg_usleep(long microsec) { usleep((useconds_t) microsec); }
void		FlushErrorState(void) {}
int			pg_strcasecmp(const char *a, const char *b) { return strcasecmp(a, b); }
int32		GetSharedNextRecordTypmod(void *s) { (void) s; return 0; }
void	   *MemoryContextAlloc(MemoryContext c, Size n) { (void) c; return palloc(n); }
void	   *MemoryContextAllocZero(MemoryContext c, Size n) { (void) c; return palloc0(n); }
void		MemoryContextDeleteImpl(MemoryContext c, const char *f, const char *fn, int l) { (void) c; (void) f; (void) fn; (void) l; }

/* one dummy context: allocations are malloc'ed and live as long as the process (a test process) */
static MemoryContextData dummy_cxt;

MemoryContext
AllocSetContextCreateInternal(MemoryContext parent, const char *name, Size a, Size b, Size c)
{
	(void) parent; (void) name; (void) a; (void) b; (void) c;
	return &dummy_cxt;
}

TupleDesc
CreateTupleDescCopy(TupleDesc d)
{
	const Size	sz = offsetof(struct TupleDescData, attrs) + d->natts * sizeof(FormData_pg_attribute);
	TupleDesc	c = palloc(sz);

	memcpy(c, d, sz);
	c->tdrefcount = -1;
	return c;
}

void		FreeTupleDesc(TupleDesc d) { pfree(d); }
/* record types never travel here: the remapper is an opaque token the layer keeps per connection */
TupleRemapper *CreateTupleRemapper(void) { return (TupleRemapper *) palloc0(8); }
void		DestroyTupleRemapper(TupleRemapper *r) { pfree(r); }
MinimalTuple TRCheckAndRemap(TupleRemapper *r, TupleDesc d, MinimalTuple t) { (void) r; (void) d; return t; }

/*
 * The layer becomes current the way a backend makes it so: through the REFERENCE's own registry (cdbmotion.c:1315
 * RegisterIPCLayerImpl, :1259 SetCurrentMotionIPCLayer).  First three stand-ins take the slots that interconnect.so's
 * tcp / udpifc / proxy implementations take on a real cluster (contrib/interconnect/ic_modules.c:150-152), then the module
 * under test registers itself exactly as its _PG_init does, and "gp_interconnect_type = cbgpu" selects it by name: the
 * registry accepts a fourth implementation without any change to core.
 */
static MotionIPCLayer stand_in[3] = {
	{.ic_type = INTERCONNECT_TYPE_TCP, .type_name = "tcp"},
	{.ic_type = INTERCONNECT_TYPE_UDPIFC, .type_name = "udpifc"},
	{.ic_type = INTERCONNECT_TYPE_PROXY, .type_name = "proxy"},
};

/* 0 = registered and selected; called once per process */
int
ref_motion_register(void)
{
	static bool done = false;

	if (done)
		return 0;
	if (setjmp(ref_jmp))
		return -1;
	for (int i = 0; i < 3; i++)
		RegisterIPCLayerImpl(&stand_in[i]);
	RegisterIPCLayerImpl(&cbgpu_ipc_layer);
	if (!CheckGpInterconnectTypeStr(&(char *) {(char *) "cbgpu"}))
		return -2;
	SetCurrentMotionIPCLayer("cbgpu");
	if (CurrentMotionIPCLayer != &cbgpu_ipc_layer || Gp_interconnect_type != (int) cbgpu_ipc_layer.ic_type)
		return -3;
	done = true;
	return 0;
}

/* BROADCAST_SEGIDX (cdb/tupchunk.h:59), for drivers that build route arrays */
int
ref_motion_broadcast_route(void)
{
	return BROADCAST_SEGIDX;
}

/* a FIFTH implementation must be refused ("There is no free entry"), and one whose type is already there too */
int
ref_motion_register_refusals(void)
{
	static MotionIPCLayer extra = {.ic_type = INTERCONNECT_TYPE_UDP2, .type_name = "extra"};
	int			refused = 0;

	if (ref_motion_register() != 0)
		return -1;
	if (setjmp(ref_jmp))
		refused++;
	else
		RegisterIPCLayerImpl(&extra);
	return refused;
}

/* InitSerTupInfo (tupser.c:70) asks the syscache for every column type's length / by-value flag / typtype: answered from
 * the column list the driver was given */
static struct
{
	Oid			oid;
	int16		len;
	bool		byval;
}			known_types[64];
static int	nknown = 0;

HeapTuple
SearchSysCache1(int cacheId, Datum key1)
{
	(void) cacheId;
	for (int i = 0; i < nknown; i++)
		if (known_types[i].oid == DatumGetObjectId(key1))
		{
			HeapTuple	t = palloc0(HEAPTUPLESIZE + MAXALIGN(SizeofHeapTupleHeader) + sizeof(FormData_pg_type));
			Form_pg_type pt;

			t->t_data = (HeapTupleHeader) ((char *) t + HEAPTUPLESIZE);
			t->t_data->t_hoff = MAXALIGN(SizeofHeapTupleHeader);
			pt = (Form_pg_type) GETSTRUCT(t);
			pt->oid = known_types[i].oid;
			pt->typlen = known_types[i].len;
			pt->typbyval = known_types[i].byval;
			pt->typtype = TYPTYPE_BASE;
			pt->typisdefined = true;
			return t;
		}
	return NULL;
}

void
ReleaseSysCache(HeapTuple t)
{
	pfree(t);
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

static EState *
make_estate(int role, int my_index, int nsenders, int nreceivers)
{
	EState	   *es = palloc0(sizeof(EState));
	SliceTable *tab = palloc0(sizeof(SliceTable));

	tab->numSlices = 2;
	tab->slices = palloc0(sizeof(ExecSlice) * 2);
	tab->localSlice = role == 0 ? 1 : 0;
	for (int k = 0; k < 2; k++)
	{
		ExecSlice  *s = &tab->slices[k];
		const int	n = k == 0 ? nreceivers : nsenders;

		s->sliceIndex = k;
		s->rootIndex = 0;
		s->parentIndex = k == 0 ? -1 : 0;
		s->children = k == 0 ? list_make1_int(1) : NIL;
		s->primaryProcesses = NIL;
		for (int i = 0; i < n; i++)
		{
			CdbProcess *p = palloc0(sizeof(CdbProcess));

			p->type = T_CdbProcess;
			p->pid = 1000 + (k == 0 ? 0 : nreceivers) + i;	/* the endpoints' "pids": one per (slice, position) */
			p->contentid = i;
			s->primaryProcesses = lappend(s->primaryProcesses, p);
		}
	}
	MyProcPid = 1000 + (role == 0 ? nreceivers : 0) + my_index;
	currentSliceId = tab->localSlice;
	es->es_sliceTable = tab;
	es->es_query_cxt = &dummy_cxt;
	return es;
}

/*
 * One process of the plan.  role 0 = sender (slice 1, position my_index), 1 = receiver (slice 0).
 * Sender: rows as ref_tupser_serialize takes them (values row-major; varlena attributes: offsets into varbuf of 4-byte
 * header datums), routes[r] = target route of row r (ref_motion_broadcast_route(): broadcast).  Returns rows sent, or -(10 + r) when SendTuple said
 * STOP_SENDING at row r.
 * Receiver: collects up to maxrows rows (values / nulls row-major, varlenas copied into varbuf_out with the header they
 * arrived with), src_out[r] = the route row r came from; stop_after >= 0: after that many rows it sends stop messages
 * (SetMotionSentinel..., the squelch path) and drains.  Returns rows received.
 * < 0: -1 an ereport(ERROR) (ref_aocs_last_error() has the text), -3 too many rows.
 */
int64
ref_motion_run(int role, int my_index, int nsenders, int nreceivers, int session, int command, int natts, const int *typid,
			   const int *typlen, const int *byval, const int *align, const int *storage, const int64 *values,
			   const unsigned char *varbuf, const unsigned char *nulls, int64 nrows, const int16 *routes, int64 *values_out,
			   unsigned char *nulls_out, int16 *src_out, int64 maxrows, unsigned char *varbuf_out, int64 varcap, int64 stop_after)
{
	EState	   *es;
	MotionLayerState *ml;
	ChunkTransportState *ts;
	TupleDesc	desc;
	Datum	   *vals;
	bool	   *isnull;
	int64		done = 0,
				vpos = 0;

	if (ref_motion_register() != 0)	/* has its own setjmp: before this function arms ref_jmp */
		return -5;
	if (setjmp(ref_jmp))
		return -1;
	gp_session_id = session;
	gp_command_count = command;
	CurrentMemoryContext = &dummy_cxt;
	TopMemoryContext = &dummy_cxt;
	es = make_estate(role, my_index, nsenders, nreceivers);
	desc = make_desc(natts, typid, typlen, byval, align, storage);
	nknown = 0;
	for (int a = 0; a < natts && nknown < 64; a++)
	{
		known_types[nknown].oid = (Oid) typid[a];
		known_types[nknown].len = (int16) typlen[a];
		known_types[nknown].byval = byval[a] != 0;
		nknown++;
	}
	CurrentMotionIPCLayer->SetupInterconnect(es);
	ts = es->interconnect_context;
	ml = createMotionLayerState(1);
	UpdateMotionLayerNode(ml, 1, false, desc);
	/* the receiver must know how many streams to expect (ExecInitMotion -> UpdateMotionExpectedReceivers, nodeMotion.c) */
	if (role == 1)
		UpdateMotionExpectedReceivers(ml, es->es_sliceTable);
	vals = palloc(sizeof(Datum) * (natts ? natts : 1));
	isnull = palloc(sizeof(bool) * (natts ? natts : 1));
	if (role == 0)
	{
		TupleTableSlot slot;

		memset(&slot, 0, sizeof(slot));
		slot.type = T_TupleTableSlot;
		*(const TupleTableSlotOps **) &slot.tts_ops = &TTSOpsVirtual;
		slot.tts_tupleDescriptor = desc;
		slot.tts_values = vals;
		slot.tts_isnull = isnull;
		slot.tts_nvalid = (AttrNumber) natts;
		for (int64 r = 0; r < nrows; r++)
		{
			for (int a = 0; a < natts; a++)
			{
				isnull[a] = nulls && nulls[r * natts + a];
				vals[a] = isnull[a] ? (Datum) 0 : typlen[a] == -1 ? PointerGetDatum(varbuf + values[r * natts + a]) : (Datum) values[r * natts + a];
			}
			if (SendTuple(ml, ts, 1, &slot, routes[r]) == STOP_SENDING)
			{
				done = -(10 + r);
				break;
			}
			done++;
		}
		SendEndOfStream(ml, ts, 1);
	}
	else
	{
		for (;;)
		{
			int16		src = ANY_ROUTE;
			MinimalTuple mt = RecvTupleFrom(ml, ts, 1, ANY_ROUTE);
			HeapTupleData htup;

			if (mt == NULL)
				break;			/* every sender's end of stream has arrived */
			(void) src;
			if (done >= maxrows)
				return -3;
			htup.t_len = mt->t_len + MINIMAL_TUPLE_OFFSET;
			htup.t_data = (HeapTupleHeader) ((char *) mt - MINIMAL_TUPLE_OFFSET);
			heap_deform_tuple(&htup, desc, vals, isnull);
			for (int a = 0; a < natts; a++)
			{
				nulls_out[done * natts + a] = isnull[a] ? 1 : 0;
				if (isnull[a])
					values_out[done * natts + a] = 0;
				else if (typlen[a] == -1)
				{
					const Size	sz = VARSIZE_ANY(DatumGetPointer(vals[a]));

					if (vpos + (int64) sz > varcap)
						return -4;
					memcpy(varbuf_out + vpos, DatumGetPointer(vals[a]), sz);
					values_out[done * natts + a] = vpos;
					vpos += (int64) sz;
				}
				else
					values_out[done * natts + a] = (int64) vals[a];
			}
			src_out[done] = -1;
			done++;
			if (stop_after >= 0 && done == stop_after)
			{
				/* the consumer has had enough (a LIMIT above the Motion: ExecSquelchMotion -> SendStopMessage) */
				SendStopMessage(ml, ts, 1);
				break;
			}
		}
	}
	EndMotionLayerNode(ml, 1, false);
	CurrentMotionIPCLayer->TeardownInterconnect(ts, false);
	return done;
}
