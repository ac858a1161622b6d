/*
 * integration/cbgpu_shim.c - the backend-side shim of INTEGRATION.md, as a real extension source.
 *
 * Compiled (type-checked) against the REFERENCE's own headers by tests/test_shim_compiles.py
 * (`gcc -fsyntax-only -I oracle/ref_shim -I /root/reference/src/include -I include`); it cannot be
 * linked or run here because the reference server cannot be built in this container (no bison /
 * flex), so it is the drop-in boundary written down precisely, not a tested component.
 *
 * What it does (SURVEY.md 8b route 2, row f1):
 *   _PG_init                       installs ExecutorStart_hook (executor/execMain.c:124)
 *   cbgpu_ExecutorStart            standard_ExecutorStart, then walks the PlanState tree; the root of every
 *                                  maximal sub-tree made only of SeqScan / Hash / HashJoin / Agg / Motion nodes
 *                                  whose expressions translate gets its ExecProcNode replaced
 *                                  (ExecSetExecProcNode, execProcnode.c:580: the GPDB wrapper with its instrumentation,
 *                                  probes and interrupt checks stays around it; nodes/execnodes.h:1065)
 *   translate_plan / translate_expr   Plan -> CbPlan, Expr -> CbExpr (include/cb_plan.h), List* -> arrays,
 *                                  operator / aggregate Oids -> CbOp / CbAggFn by catalog name
 *   cbgpu_ExecNode                 ExecProcNodeMtd: pulls a CbTupleTableSlot from cb_ExecProcNode and stores it as
 *                                  a virtual tuple (ExecStoreVirtualTuple, as aocsam_handler.c:768 does);
 *                                  library errors become ereport(ERROR) only after the device state is released
 *   memory context reset callback  frees device state when the query context goes away on an error path
 *
 * Loading the range table is integration/cbgpu_shim_storage.c (cbgpu_shim_load_relation: only the columns a scan's
 * target list and quals mention are read, so Var.varattno is remapped to the position among them); shipping the NCCL
 * rendezvous token with the dispatched plan is sketched in INTEGRATION.md and left as an extern hook
 * (cbgpu_shim_interconnect), so this file stays about the operator boundary.
 *
 * character(n > 1) / varchar / text columns live on the device as codes of a per-segment dictionary the loader builds
 * (cbgpu_shim_dict).  Codes mean nothing outside the segment that made them, so such a column is accepted only where
 * it never leaves its scan: in a SeqScan qual of the form column = 'literal' / column <> 'literal', the literal being
 * translated to the segment's code once the relation is loaded.  Anything else on such a column declines the sub-tree.
 */
#include "postgres.h"

#include "access/tupdesc.h"
#include "access/xact.h"
#include "catalog/pg_type.h"
#include "cdb/cdbutil.h"
#include "cdb/cdbvars.h"
#include "executor/executor.h"
#include "executor/tuptable.h"
#include "fmgr.h"
#include "miscadmin.h"
#include "nodes/execnodes.h"
#include "nodes/extensible.h"
#include "nodes/makefuncs.h"
#include "optimizer/planner.h"
#include "utils/guc.h"
#include "nodes/nodeFuncs.h"
#include "nodes/plannodes.h"
#include "nodes/primnodes.h"
#include "optimizer/optimizer.h"
#include "access/sysattr.h"
#include "utils/builtins.h"
#include "utils/lsyscache.h"
#include "utils/memutils.h"
#include "utils/numeric.h"
#include "utils/rel.h"

#include "cb_exec.h"

PG_MODULE_MAGIC;

void		_PG_init(void);

/* provided by the loader / dispatcher glue */
extern cbgpu_rel *cbgpu_shim_load_relation(cbgpu_ctx *ctx, Relation rel, List *projected_attnos);
extern cbgpu_dict *cbgpu_shim_dict(Oid relid, AttrNumber attno);

/* one SeqScan of the sub-tree: its range-table index and the columns it needs, ascending (= device column order) */
typedef struct ShimScan
{
	Index		scanrelid;
	List	   *attnos;
} ShimScan;

/* a string literal compared with a dictionary column in a scan qual: becomes a code after the relation is loaded */
typedef struct ShimPendingConst
{
	CbExpr	   *x;
	Index		scanrelid;
	AttrNumber	attno;
	Const	   *c;
} ShimPendingConst;

/* translation context (one sub-tree at a time, backend-local) */
static ShimScan *cur_scan = NULL;		/* the SeqScan whose expressions are being translated, else NULL   */
static bool cur_in_qual = false;
static List *pending_consts = NIL;
/* integration/cbgpu_shim_interconnect.c: the session's NCCL communicator, its token shipped as a synced GUC */
extern CbInterconnect *cbgpu_shim_interconnect(cbgpu_ctx *ctx, EState *estate);
extern void cbgpu_shim_define_gucs(void);
extern void cbgpu_shim_qd_prepare(bool previous_query_failed);

static ExecutorStart_hook_type prev_ExecutorStart = NULL;
static ExecutorEnd_hook_type prev_ExecutorEnd = NULL;
static cbgpu_ctx *shim_ctx = NULL;		/* one device context per backend, created after fork, on first use */

typedef struct CbgpuShim
{
	CbEState   *cbestate;
	CbPlanState *cbps;
	MemoryContextCallback reset_cb;
	ExecProcNodeMtd saved_ExecProcNode;
	int			natts;
	CbTypeId   *atttypes;
	int32	   *attdscales;
	/* columns that leave as serialised partial aggregate states (the sub-tree's top is a Partial Aggregate whose Finalize
	 * stage stays on the CPU): 0 plain value, 1 numeric_avg_serialize form, 2 int8_avg_serialize form */
	int		   *attserial;
	cbgpu_rel **rt;				/* the base tables loaded for this sub-tree: device memory this shim owns */
	int			nrt;
} CbgpuShim;

/* ------------------------------------------------------------------------------------------
 * types
 * ------------------------------------------------------------------------------------------ */
static bool
translate_type(Oid typid, int32 typmod, CbTypeId *out, int32 *dscale)
{
	*dscale = 0;
	switch (typid)
	{
		case INT4OID: *out = CB_INT4; return true;
		case INT8OID: *out = CB_INT8; return true;
		case DATEOID: *out = CB_DATE; return true;
		case FLOAT8OID: *out = CB_FLOAT8; return true;
		case BOOLOID: *out = CB_BOOL; return true;
		case NUMERICOID:
			/* numeric(p,s): typmod = ((p << 16) | s) + VARHDRSZ (utils/adt/numeric.c numerictypmodin) */
			if (typmod < (int32) VARHDRSZ)
				return false;	/* unconstrained numeric has no fixed scale */
			*out = CB_NUMERIC;
			*dscale = (typmod - VARHDRSZ) & 0xffff;
			return true;
		case BPCHAROID:
			if (typmod != (int32) VARHDRSZ + 1)
				return false;	/* character(n > 1) / varchar need the loader's dictionary: not translated here */
			*out = CB_BPCHAR1;
			return true;
		default:
			return false;
	}
}

/* ------------------------------------------------------------------------------------------
 * expressions (nodes/primnodes.h) -> CbExpr
 * ------------------------------------------------------------------------------------------ */
static CbExpr *translate_expr(Expr *e);

static CbExpr *
new_expr(CbNodeTag tag, CbTypeId t, int32 dscale)
{
	CbExpr	   *x = (CbExpr *) palloc0(sizeof(CbExpr));

	x->tag = tag;
	x->restype = t;
	x->dscale = dscale;
	return x;
}

static bool
translate_args(CbExpr *x, List *args)
{
	ListCell   *lc;
	int			i = 0;

	x->nargs = list_length(args);
	x->args = (CbExpr **) palloc0(sizeof(CbExpr *) * Max(x->nargs, 1));
	foreach(lc, args)
	{
		Node	   *a = (Node *) lfirst(lc);

		if (IsA(a, TargetEntry))
			a = (Node *) ((TargetEntry *) a)->expr;		/* Aggref.args is a list of TargetEntry */
		x->args[i] = translate_expr((Expr *) a);
		if (x->args[i] == NULL)
			return false;
		i++;
	}
	return true;
}

static bool
is_string_type(Oid typid, int32 typmod)
{
	return typid == VARCHAROID || typid == TEXTOID || (typid == BPCHAROID && typmod != (int32) VARHDRSZ + 1);
}

static Node *
strip_relabel(Node *n)
{
	while (n && IsA(n, RelabelType))
		n = (Node *) ((RelabelType *) n)->arg;
	return n;
}

/* 1-based position of v in an integer list, 0 when absent */
static int
list_position_int(List *l, int v)
{
	ListCell   *lc;
	int			pos = 0;

	foreach(lc, l)
	{
		pos++;
		if (lfirst_int(lc) == v)
			return pos;
	}
	return 0;
}

static CbExpr *
translate_expr(Expr *e)
{
	CbTypeId	t;
	int32		ds;

	if (e == NULL)
		return NULL;
	switch (nodeTag(e))
	{
		case T_Var:
			{
				Var		   *v = (Var *) e;
				CbExpr	   *x;

				if (v->varlevelsup != 0 || v->varattno <= 0)
					return NULL;
				if (is_string_type(v->vartype, v->vartypmod))
				{
					/* character(n) / varchar / text columns are dictionary codes on the device (the loader's choice:
					 * integration/cbgpu_shim_storage.c gives them CB_DICT32), here and in every node above */
					t = CB_DICT32;
					ds = 0;
				}
				else if ((v->vartype == BYTEAOID || v->vartype == INTERNALOID || (v->vartype == NUMERICOID && v->vartypmod < (int32) VARHDRSZ)) &&
						 (v->varno == OUTER_VAR || v->varno == INNER_VAR))
				{
					/* a partial aggregate's serialised state travelling up (Motion target lists, the Finalize Aggref's
					 * argument), or an aggregate's / expression's numeric result (no typmod above the node that computes
					 * it): the device keeps states as (N, sum) columns and numerics as scaled integers whose type and
					 * display scale are the producing expression's, known once the child is translated - resolve_plan
					 * fills them in */
					t = (CbTypeId) 0;
					ds = 0;
				}
				else if (!translate_type(v->vartype, v->vartypmod, &t, &ds))
					return NULL;
				x = new_expr(T_CbVar, t, ds);
				x->varno = v->varno;	/* INNER_VAR / OUTER_VAR keep their values (65000 / 65001) */
				x->varattno = v->varattno;
				if (cur_scan && v->varno == cur_scan->scanrelid)
				{
					/* the device relation holds only the projected columns, in attribute order */
					ListCell   *lc;
					int			pos = 0;

					x->varattno = 0;
					foreach(lc, cur_scan->attnos)
					{
						pos++;
						if (lfirst_int(lc) == v->varattno)
							x->varattno = pos;
					}
					if (x->varattno == 0)
						return NULL;
				}
				return x;
			}
		case T_Const:
			{
				Const	   *c = (Const *) e;
				CbExpr	   *x;

				if (!translate_type(c->consttype, c->consttypmod, &t, &ds) && c->consttype != NUMERICOID)
					return NULL;
				if (c->consttype == NUMERICOID)
				{
					/* a numeric literal: its own display scale, value scaled to int64 */
					Numeric		n = DatumGetNumeric(c->constvalue);
					char	   *s;

					if (c->constisnull || numeric_is_nan(n))
						return NULL;
					t = CB_NUMERIC;
					ds = (int32) DatumGetInt32(DirectFunctionCall1(numeric_scale, c->constvalue));
					s = DatumGetCString(DirectFunctionCall1(numeric_out, c->constvalue));
					x = new_expr(T_CbConst, t, ds);
					{
						/* digits without the point = the value scaled by 10^ds */
						int64		v = 0;
						bool		neg = false;

						for (char *p = s; *p; p++)
						{
							if (*p == '-')
								neg = true;
							else if (*p >= '0' && *p <= '9')
							{
								if (v > (PG_INT64_MAX - 9) / 10)
									return NULL;
								v = v * 10 + (*p - '0');
							}
						}
						x->constval = neg ? -v : v;
					}
					return x;
				}
				x = new_expr(T_CbConst, t, ds);
				x->constisnull = c->constisnull;
				if (!c->constisnull)
				{
					switch (t)
					{
						case CB_INT4: x->constval = DatumGetInt32(c->constvalue); break;
						case CB_DATE: x->constval = DatumGetInt32(c->constvalue); break;	/* DateADT */
						case CB_INT8: x->constval = DatumGetInt64(c->constvalue); break;
						case CB_BOOL: x->constval = DatumGetBool(c->constvalue); break;
						case CB_FLOAT8: memcpy(&x->constval, &c->constvalue, sizeof(int64)); break;
						case CB_BPCHAR1: x->constval = (unsigned char) *VARDATA_ANY(DatumGetPointer(c->constvalue)); break;
						default: return NULL;
					}
				}
				return x;
			}
		case T_OpExpr:
			{
				OpExpr	   *o = (OpExpr *) e;
				char	   *name = get_opname(o->opno);
				CbExpr	   *x;
				int			op;

				if (name == NULL || list_length(o->args) != 2)
					return NULL;
				{
					/* dictionary column = / <> string literal, in the qual of the scan that owns the column */
					Node	   *l = strip_relabel((Node *) linitial(o->args));
					Node	   *r = strip_relabel((Node *) lsecond(o->args));

					if (IsA(l, Const) && IsA(r, Var))
					{
						Node	   *tmp = l;

						l = r;
						r = tmp;
					}
					if (IsA(l, Var) && is_string_type(((Var *) l)->vartype, ((Var *) l)->vartypmod))
					{
						Var		   *v = (Var *) l;
						ShimPendingConst *pc;
						CbExpr	   *cx;

						if (!cur_scan || !cur_in_qual || v->varno != cur_scan->scanrelid || !IsA(r, Const) || ((Const *) r)->constisnull ||
							!is_string_type(((Const *) r)->consttype, -1) || (strcmp(name, "=") != 0 && strcmp(name, "<>") != 0))
							return NULL;
						x = new_expr(T_CbOpExpr, CB_BOOL, 0);
						x->op = strcmp(name, "=") == 0 ? CB_OP_EQ : CB_OP_NE;
						x->nargs = 2;
						x->args = (CbExpr **) palloc0(sizeof(CbExpr *) * 2);
						x->args[0] = new_expr(T_CbVar, CB_DICT32, 0);
						x->args[0]->varno = v->varno;
						x->args[0]->varattno = list_position_int(cur_scan->attnos, v->varattno);
						cx = new_expr(T_CbConst, CB_DICT32, 0);
						cx->constval = -1;		/* no row has this code: "the column never holds the value" */
						x->args[1] = cx;
						if (x->args[0]->varattno == 0)
							return NULL;
						pc = (ShimPendingConst *) palloc(sizeof(ShimPendingConst));
						pc->x = cx;
						pc->scanrelid = cur_scan->scanrelid;
						pc->attno = v->varattno;
						pc->c = (Const *) r;
						pending_consts = lappend(pending_consts, pc);
						return x;
					}
				}
				if (strcmp(name, "+") == 0) op = CB_OP_ADD;
				else if (strcmp(name, "-") == 0) op = CB_OP_SUB;
				else if (strcmp(name, "*") == 0) op = CB_OP_MUL;
				else if (strcmp(name, "=") == 0) op = CB_OP_EQ;
				else if (strcmp(name, "<>") == 0) op = CB_OP_NE;
				else if (strcmp(name, "<") == 0) op = CB_OP_LT;
				else if (strcmp(name, "<=") == 0) op = CB_OP_LE;
				else if (strcmp(name, ">") == 0) op = CB_OP_GT;
				else if (strcmp(name, ">=") == 0) op = CB_OP_GE;
				else
					return NULL;
				x = new_expr(T_CbOpExpr, CB_BOOL, 0);
				x->op = op;
				if (!translate_args(x, o->args))
					return NULL;
				if (op < CB_OP_EQ)
				{
					/* result type and display scale as numeric_add / numeric_mul keep them (utils/adt/numeric.c:2491,2645) */
					CbExpr	   *a = x->args[0], *b = x->args[1];

					if (!translate_type(o->opresulttype, -1, &t, &ds) && o->opresulttype != NUMERICOID)
						return NULL;
					x->restype = o->opresulttype == NUMERICOID ? CB_NUMERIC : t;
					x->dscale = op == CB_OP_MUL ? a->dscale + b->dscale : Max(a->dscale, b->dscale);
				}
				return x;
			}
		case T_BoolExpr:
			{
				BoolExpr   *b = (BoolExpr *) e;
				CbExpr	   *x = new_expr(T_CbBoolExpr, CB_BOOL, 0);

				x->op = b->boolop == AND_EXPR ? CB_AND_EXPR : b->boolop == OR_EXPR ? CB_OR_EXPR : CB_NOT_EXPR;
				return translate_args(x, b->args) ? x : NULL;
			}
		case T_Aggref:
			{
				Aggref	   *a = (Aggref *) e;
				char	   *name = get_func_name(a->aggfnoid);
				CbExpr	   *x;
				int			fn;

				if (name == NULL || a->aggdistinct != NIL || a->aggorder != NIL || a->aggfilter != NULL || a->aggdirectargs != NIL)
					return NULL;
				if (strcmp(name, "count") == 0) fn = a->aggstar ? CB_AGG_COUNT_STAR : CB_AGG_COUNT;
				else if (strcmp(name, "sum") == 0) fn = CB_AGG_SUM;
				else if (strcmp(name, "avg") == 0) fn = CB_AGG_AVG;
				else if (strcmp(name, "min") == 0) fn = CB_AGG_MIN;
				else if (strcmp(name, "max") == 0) fn = CB_AGG_MAX;
				else
					return NULL;
				{
					/* a partial aggregate's aggtype is its serialised transition type (bytea for the numeric aggregates:
					 * mark_partial_aggref, optimizer/plan/planner.c); the device keeps typed (N, sum) states, so the type
					 * that matters is the aggregate's own result type */
					Oid			rettype = a->aggtype;

					if (DO_AGGSPLIT_SKIPFINAL(a->aggsplit))
						rettype = get_func_rettype(a->aggfnoid);
					if (!translate_type(rettype, -1, &t, &ds) && rettype != NUMERICOID)
						return NULL;
					x = new_expr(T_CbAggref, rettype == NUMERICOID ? CB_NUMERIC : t, 0);
				}
				x->op = fn;
				if (!translate_args(x, a->args))
					return NULL;
				if (x->nargs == 1 && x->restype == CB_NUMERIC)
					x->dscale = x->args[0]->dscale;		/* numeric_sum keeps the input display scale (numeric.c:6091) */
				return x;
			}
		case T_RelabelType:
			return translate_expr(((RelabelType *) e)->arg);
		default:
			return NULL;
	}
}

static bool
translate_exprs(List *l, int32 *n, CbExpr ***out)
{
	ListCell   *lc;
	int			i = 0;

	*n = list_length(l);
	*out = (CbExpr **) palloc0(sizeof(CbExpr *) * Max(*n, 1));
	foreach(lc, l)
	{
		(*out)[i] = translate_expr((Expr *) lfirst(lc));
		if ((*out)[i] == NULL)
			return false;
		i++;
	}
	return true;
}

/* ------------------------------------------------------------------------------------------
 * plan nodes (nodes/plannodes.h) -> CbPlan
 * ------------------------------------------------------------------------------------------ */
static CbPlan *translate_plan(Plan *p, EState *estate, List **rels);

static bool
fill_plan(CbPlan *c, CbNodeTag tag, Plan *p)
{
	ListCell   *lc;
	int			i = 0;

	c->type = tag;
	c->plan_node_id = p->plan_node_id;
	c->plan_rows = p->plan_rows;
	c->ntargets = list_length(p->targetlist);
	c->targetlist = (CbTargetEntry *) palloc0(sizeof(CbTargetEntry) * Max(c->ntargets, 1));
	foreach(lc, p->targetlist)
	{
		TargetEntry *te = (TargetEntry *) lfirst(lc);

		c->targetlist[i].expr = translate_expr(te->expr);
		if (c->targetlist[i].expr == NULL)
			return false;
		c->targetlist[i].resno = te->resno;
		c->targetlist[i].resname = te->resname;
		i++;
	}
	{
		bool		ok;

		cur_in_qual = true;
		ok = translate_exprs(p->qual, &c->nquals, &c->qual);
		cur_in_qual = false;
		return ok;
	}
}

static CbPlan *
translate_plan(Plan *p, EState *estate, List **rels)
{
	if (p == NULL)
		return NULL;
	switch (nodeTag(p))
	{
		case T_SeqScan:
			{
				CbSeqScan  *c = (CbSeqScan *) palloc0(sizeof(CbSeqScan));
				ShimScan   *sc = (ShimScan *) palloc0(sizeof(ShimScan));
				Bitmapset  *used = NULL;
				int			k = -1;
				bool		ok;

				/* the columns this scan reads: what its target list and quals mention (as aoco_beginscan_extractcolumns
				 * projects, access/aocs/aocsam_handler.c:612) */
				sc->scanrelid = ((Scan *) p)->scanrelid;
				pull_varattnos((Node *) p->targetlist, sc->scanrelid, &used);
				pull_varattnos((Node *) p->qual, sc->scanrelid, &used);
				while ((k = bms_next_member(used, k)) >= 0)
				{
					const int	attno = k + FirstLowInvalidHeapAttributeNumber;

					if (attno <= 0)
						return NULL;	/* system columns, whole-row references */
					sc->attnos = lappend_int(sc->attnos, attno);
				}
				if (sc->attnos == NIL)
					return NULL;
				cur_scan = sc;
				ok = fill_plan(&c->plan, T_CbSeqScan, p);
				cur_scan = NULL;
				if (!ok)
					return NULL;
				/* the range-table index of this scan in the GPU executor = its position in `rels` */
				*rels = lappend(*rels, sc);
				c->scanrelid = list_length(*rels);
				return &c->plan;
			}
		case T_Hash:
			{
				CbHash	   *c = (CbHash *) palloc0(sizeof(CbHash));

				if (!fill_plan(&c->plan, T_CbHash, p) || !translate_exprs(((Hash *) p)->hashkeys, &c->nhashkeys, &c->hashkeys))
					return NULL;
				c->plan.lefttree = translate_plan(outerPlan(p), estate, rels);
				return c->plan.lefttree ? &c->plan : NULL;
			}
		case T_HashJoin:
			{
				HashJoin   *hj = (HashJoin *) p;
				CbHashJoin *c = (CbHashJoin *) palloc0(sizeof(CbHashJoin));

				if (!fill_plan(&c->plan, T_CbHashJoin, p))
					return NULL;
				switch (hj->join.jointype)
				{
					case JOIN_INNER: c->jointype = CB_JOIN_INNER; break;
					case JOIN_LEFT: c->jointype = CB_JOIN_LEFT; break;
					case JOIN_SEMI: c->jointype = CB_JOIN_SEMI; break;
					case JOIN_ANTI: c->jointype = CB_JOIN_ANTI; break;
					case JOIN_RIGHT: c->jointype = CB_JOIN_RIGHT; break;
					case JOIN_FULL: c->jointype = CB_JOIN_FULL; break;
					case JOIN_LASJ_NOTIN: c->jointype = CB_JOIN_LASJ_NOTIN; break;
					default: return NULL;	/* unique-ified and dedup semi joins stay on the CPU */
				}
				if (!translate_exprs(hj->hashkeys, &c->nhashkeys, &c->hashkeys) ||
					!translate_exprs(hj->join.joinqual, &c->njoinquals, &c->joinqual))
					return NULL;
				c->plan.lefttree = translate_plan(outerPlan(p), estate, rels);
				c->plan.righttree = translate_plan(innerPlan(p), estate, rels);
				return (c->plan.lefttree && c->plan.righttree) ? &c->plan : NULL;
			}
		case T_Agg:
			{
				Agg		   *a = (Agg *) p;
				CbAgg	   *c = (CbAgg *) palloc0(sizeof(CbAgg));

				if (a->groupingSets != NIL || a->chain != NIL || (a->aggstrategy != AGG_HASHED && a->aggstrategy != AGG_PLAIN) ||
					!fill_plan(&c->plan, T_CbAgg, p))
					return NULL;
				c->aggstrategy = a->aggstrategy == AGG_HASHED ? CB_AGG_HASHED : CB_AGG_PLAIN;
				if (a->aggsplit == AGGSPLIT_SIMPLE)
					c->aggsplit = CB_AGGSPLIT_SIMPLE;
				else if (a->aggsplit == AGGSPLIT_INITIAL_SERIAL)
					c->aggsplit = CB_AGGSPLIT_INITIAL_SERIAL;
				else if (a->aggsplit == AGGSPLIT_FINAL_DESERIAL)
					c->aggsplit = CB_AGGSPLIT_FINAL_DESERIAL;
				else
					return NULL;
				c->numCols = a->numCols;
				c->grpColIdx = (int32_t *) palloc0(sizeof(int32_t) * Max(a->numCols, 1));
				for (int i = 0; i < a->numCols; i++)
					c->grpColIdx[i] = a->grpColIdx[i];
				c->numGroups = a->numGroups;
				c->streaming = a->streaming;
				c->plan.lefttree = translate_plan(outerPlan(p), estate, rels);
				return c->plan.lefttree ? &c->plan : NULL;
			}
		case T_Motion:
			{
				Motion	   *m = (Motion *) p;
				CbMotion   *c = (CbMotion *) palloc0(sizeof(CbMotion));

				if (!fill_plan(&c->plan, T_CbMotion, p))
					return NULL;
				if (m->sendSorted)
				{
					/* merge receive (execMotionSortedReceiver, nodeMotion.c:433): the keys as (column, descending); only the
					 * default NULL placement (last ascending, first descending) is what the device comparator implements */
					if (m->motionType != MOTIONTYPE_GATHER && m->motionType != MOTIONTYPE_GATHER_SINGLE)
						return NULL;
					c->nsortkeys = m->numSortCols;
					c->sortkeys = (CbSortKey *) palloc0(sizeof(CbSortKey) * Max(m->numSortCols, 1));
					for (int i = 0; i < m->numSortCols; i++)
					{
						Oid			opfamily,
									opcintype;
						int16		strategy;

						if (!get_ordering_op_properties(m->sortOperators[i], &opfamily, &opcintype, &strategy))
							return NULL;
						c->sortkeys[i].attno = m->sortColIdx[i];
						c->sortkeys[i].descending = strategy == BTGreaterStrategyNumber;
						if (m->nullsFirst[i] != c->sortkeys[i].descending)
							return NULL;
					}
				}
				switch (m->motionType)
				{
					case MOTIONTYPE_GATHER: c->motionType = CB_MOTIONTYPE_GATHER; break;
					case MOTIONTYPE_GATHER_SINGLE: c->motionType = CB_MOTIONTYPE_GATHER_SINGLE; break;
					case MOTIONTYPE_HASH: c->motionType = CB_MOTIONTYPE_HASH; break;
					case MOTIONTYPE_BROADCAST: c->motionType = CB_MOTIONTYPE_BROADCAST; break;
					default: return NULL;
				}
				c->motionID = m->motionID;
				c->numHashSegments = m->numHashSegments;
				if (!translate_exprs(m->hashExprs, &c->nhashExprs, &c->hashExprs))
					return NULL;
				c->plan.lefttree = translate_plan(outerPlan(p), estate, rels);
				return c->plan.lefttree ? &c->plan : NULL;
			}
		default:
			return NULL;
	}
}

/* Second pass: Vars that carry a partial aggregate's state take the type and display scale of the target entry they
 * point at (the child's Aggref), bottom-up; a Finalize Aggref over such a Var then gets its display scale
 * (numeric_sum / numeric_avg keep their input's, utils/adt/numeric.c:6091). */
static bool
resolve_expr(CbExpr *x, const CbPlan *outer, const CbPlan *inner)
{
	if (x == NULL)
		return true;
	if (x->tag == T_CbVar && x->restype == 0)
	{
		const CbPlan *src = x->varno == OUTER_VAR ? outer : x->varno == INNER_VAR ? inner : NULL;
		const CbExpr *t;

		if (src == NULL || x->varattno < 1 || x->varattno > src->ntargets)
			return false;
		t = src->targetlist[x->varattno - 1].expr;
		if (t->restype == 0)
			return false;
		x->restype = t->restype;
		x->dscale = t->dscale;
	}
	for (int i = 0; i < x->nargs; i++)
		if (!resolve_expr(x->args[i], outer, inner))
			return false;
	if (x->tag == T_CbAggref && x->nargs == 1 && x->restype == CB_NUMERIC)
		x->dscale = x->args[0]->dscale;
	return true;
}

static bool
resolve_exprs(CbExpr **l, int n, const CbPlan *outer, const CbPlan *inner)
{
	for (int i = 0; i < n; i++)
		if (!resolve_expr(l[i], outer, inner))
			return false;
	return true;
}

static bool
resolve_plan(CbPlan *c)
{
	const CbPlan *outer,
			   *inner;

	if (c == NULL)
		return true;
	if (!resolve_plan(c->lefttree) || !resolve_plan(c->righttree))
		return false;
	outer = c->lefttree;
	inner = c->righttree;
	for (int i = 0; i < c->ntargets; i++)
		if (!resolve_expr(c->targetlist[i].expr, outer, inner))
			return false;
	if (!resolve_exprs(c->qual, c->nquals, outer, inner))
		return false;
	switch (c->type)
	{
		case T_CbHash:
			return resolve_exprs(((CbHash *) c)->hashkeys, ((CbHash *) c)->nhashkeys, outer, inner);
		case T_CbHashJoin:
			return resolve_exprs(((CbHashJoin *) c)->hashkeys, ((CbHashJoin *) c)->nhashkeys, outer, inner) &&
				resolve_exprs(((CbHashJoin *) c)->joinqual, ((CbHashJoin *) c)->njoinquals, outer, inner);
		case T_CbMotion:
			return resolve_exprs(((CbMotion *) c)->hashExprs, ((CbMotion *) c)->nhashExprs, outer, inner);
		default:
			return true;
	}
}

/* ------------------------------------------------------------------------------------------
 * the replaced ExecProcNode
 * ------------------------------------------------------------------------------------------ */
/* PlanState -> shim: a short list in the query context (a core patch would add one pointer to PlanState) */
typedef struct ShimEntry
{
	PlanState  *ps;
	CbgpuShim  *shim;
	struct ShimEntry *next;
} ShimEntry;
static ShimEntry *shim_list = NULL;

static CbgpuShim *
shim_of(PlanState *ps)
{
	for (ShimEntry *e = shim_list; e; e = e->next)
		if (e->ps == ps)
			return e->shim;
	elog(ERROR, "cbgpu: no shim registered for plan node %d", ps->plan->plan_node_id);
	return NULL;
}

static void
shim_release(void *arg)
{
	CbgpuShim  *shim = (CbgpuShim *) arg;

	if (shim->cbps)
		cb_ExecEndNode(shim->cbps);
	shim->cbps = NULL;
	if (shim->cbestate)
		cb_FreeExecutorState(shim->cbestate);
	shim->cbestate = NULL;
	/* cb_FreeExecutorState frees the range table ARRAY only: the relations are this shim's (a session-lived backend
	 * would otherwise lose a table's worth of HBM per query) */
	for (int i = 0; i < shim->nrt; i++)
		if (shim->rt && shim->rt[i])
			cbgpu_rel_free(shim->rt[i]);
	shim->nrt = 0;
	shim_list = NULL;			/* the entries live in the query context that is going away */
}

/* does the sub-tree exchange rows with other segments?  Then every segment must run it on the same side (GPU or CPU):
 * the take-over decision may depend on the plan's shape only, and whatever fails on one segment afterwards is an
 * ERROR for the query, never a quiet return to the CPU executor while the peers wait in the GPU interconnect */
static bool
plan_has_motion(Plan *plan)
{
	if (plan == NULL)
		return false;
	if (IsA(plan, Motion))
		return true;
	return plan_has_motion(outerPlan(plan)) || plan_has_motion(innerPlan(plan));
}

#define SHIM_LOCAL_FAILURE(has_motion, shim, ...) \
	do { \
		if (shim) \
			shim_release(shim); \
		if (has_motion) \
			ereport(ERROR, (errcode(ERRCODE_INTERNAL_ERROR), errmsg(__VA_ARGS__))); \
		return NULL; \
	} while (0)

/* the next row of the device sub-tree into `slot` (empty at end of data): shared by both routes */
static TupleTableSlot *
shim_next_tuple(CbgpuShim *shim, TupleTableSlot *slot)
{
	CbTupleTableSlot *cs;

	CHECK_FOR_INTERRUPTS();				/* miscadmin.h */
	cs = cb_ExecProcNode(shim->cbps);
	if (shim->cbestate->es_errcode)
	{
		char		msg[512];

		strlcpy(msg, cb_estate_error(shim->cbestate), sizeof(msg));
		shim_release(shim);				/* never longjmp through CUDA / NCCL frames holding device memory */
		ereport(ERROR, (errcode(ERRCODE_INTERNAL_ERROR), errmsg("cbgpu: %s", msg)));
	}
	ExecClearTuple(slot);
	if (cs == NULL || cs->tts_empty)
		return slot;					/* end of data: an empty slot (TupIsNull, executor/tuptable.h) */
	for (int i = 0; i < shim->natts; i++)
	{
		if (shim->attserial[i] != 0)
		{
			/* (N, sum) -> the serialisation function's bytea; a state is never NULL once the group exists */
			uint8		buf[256];
			const int	len = shim->attserial[i] == 1
				? cb_numeric_avg_serialize(cs->tts_state_n[i], cs->tts_state_lo[i], cs->tts_state_hi[i], shim->attdscales[i], buf, (int32) sizeof(buf))
				: cb_int8_avg_serialize(cs->tts_state_n[i], cs->tts_state_lo[i], cs->tts_state_hi[i], buf, (int32) sizeof(buf));
			bytea	   *b;

			if (len < 0)
				ereport(ERROR, (errcode(ERRCODE_INTERNAL_ERROR), errmsg("cbgpu: a partial aggregate state does not fit its serialised form")));
			b = (bytea *) palloc(VARHDRSZ + len);
			SET_VARSIZE(b, VARHDRSZ + len);
			memcpy(VARDATA(b), buf, len);
			slot->tts_isnull[i] = false;
			slot->tts_values[i] = PointerGetDatum(b);
			continue;
		}
		slot->tts_isnull[i] = cb_slot_isnull(cs, i);
		if (slot->tts_isnull[i])
		{
			slot->tts_values[i] = (Datum) 0;
			continue;
		}
		switch (shim->atttypes[i])
		{
			case CB_INT4: case CB_DATE: slot->tts_values[i] = Int32GetDatum((int32) cb_slot_int64(cs, i)); break;
			case CB_INT8: slot->tts_values[i] = Int64GetDatum(cb_slot_int64(cs, i)); break;
			case CB_BOOL: slot->tts_values[i] = BoolGetDatum(cb_slot_int64(cs, i) != 0); break;
			case CB_FLOAT8: slot->tts_values[i] = Float8GetDatum(cb_slot_float8(cs, i)); break;
			case CB_NUMERIC: case CB_NUMERIC128:
				{
					/* the library finalises numerics to the reference's digits as text (cb_numeric.c) */
					char		buf[160];

					cb_slot_text(cs, i, buf, sizeof(buf));
					slot->tts_values[i] = DirectFunctionCall3(numeric_in, CStringGetDatum(buf), ObjectIdGetDatum(InvalidOid), Int32GetDatum(-1));
					break;
				}
			default:
				{
					char		c = (char) cb_slot_int64(cs, i);

					slot->tts_values[i] = DirectFunctionCall3(bpcharin, CStringGetDatum(psprintf("%c", c)), ObjectIdGetDatum(InvalidOid),
															  Int32GetDatum(VARHDRSZ + 1));
					break;
				}
		}
	}
	return ExecStoreVirtualTuple(slot);
}

static TupleTableSlot *
cbgpu_ExecNode(PlanState *ps)			/* ExecProcNodeMtd, nodes/execnodes.h:1056 */
{
	return shim_next_tuple(shim_of(ps), ps->ps_ResultTupleSlot);
}

/* everything a sub-tree needs before its first row: translation, device context, executor state, base tables on the device,
 * string literals as dictionary codes, interconnect, the result's column types.  NULL: the sub-tree stays with the CPU
 * executor (or, with Motions below and a local failure, the query ends in an error: see SHIM_LOCAL_FAILURE) */
static CbgpuShim *
shim_prepare(Plan *plan, TupleDesc desc, EState *estate)
{
	List	   *rels = NIL;
	CbPlan	   *cplan = (pending_consts = NIL, translate_plan(plan, estate, &rels));
	CbgpuShim  *shim;
	ListCell   *lc;
	int			i = 0;
	const bool	has_motion = plan_has_motion(plan);

	/* 1. decisions that follow from the plan alone (the same on every segment): translation, node support */
	if (cplan == NULL || desc == NULL || !resolve_plan(cplan))
		return NULL;
	if (shim_ctx == NULL && cbgpu_ctx_create(GpIdentity.segindex >= 0 ? GpIdentity.segindex % Max(cbgpu_device_count(), 1) : 0, &shim_ctx) != CBGPU_OK)
		SHIM_LOCAL_FAILURE(has_motion, (CbgpuShim *) NULL, "cbgpu: no CUDA context on segment %d", GpIdentity.segindex);
	shim = (CbgpuShim *) MemoryContextAllocZero(estate->es_query_cxt, sizeof(CbgpuShim));
	shim->rt = (cbgpu_rel **) MemoryContextAllocZero(estate->es_query_cxt, sizeof(cbgpu_rel *) * Max(list_length(rels), 1));
	shim->cbestate = cb_CreateExecutorState(shim_ctx, shim->rt, list_length(rels));
	shim->cbestate->es_segindex = GpIdentity.segindex;
	shim->cbestate->es_numsegments = getgpsegmentCount();
	shim->cbps = cb_ExecInitNode(cplan, shim->cbestate, 0);	/* validates the node types before a byte is loaded */
	if (shim->cbps == NULL)
	{
		shim_release(shim);
		return NULL;
	}
	/* 2. from here on this segment is committed: load the base tables (device memory the shim owns from now on) */
	foreach(lc, rels)
	{
		ShimScan   *sc = (ShimScan *) lfirst(lc);
		Relation	r = ExecGetRangeTableRelation(estate, sc->scanrelid);

		shim->rt[i] = cbgpu_shim_load_relation(shim_ctx, r, sc->attnos);
		shim->nrt = i + 1;
		if (shim->rt[i] == NULL)
			SHIM_LOCAL_FAILURE(has_motion, shim, "cbgpu: relation \"%s\" cannot be loaded on segment %d", RelationGetRelationName(r), GpIdentity.segindex);
		shim->cbestate->es_range_table[i] = shim->rt[i];
		i++;
	}
	/* string literals -> this segment's dictionary codes, now that the dictionaries exist */
	foreach(lc, pending_consts)
	{
		ShimPendingConst *pc = (ShimPendingConst *) lfirst(lc);
		cbgpu_dict *dict = cbgpu_shim_dict(exec_rt_fetch(pc->scanrelid, estate)->relid, pc->attno);
		struct varlena *v = (struct varlena *) DatumGetPointer(pc->c->constvalue);

		if (dict == NULL)
			SHIM_LOCAL_FAILURE(has_motion, shim, "cbgpu: no dictionary for a string qual on segment %d", GpIdentity.segindex);
		/* -1 (the value occurs nowhere in this segment's files) stays -1: = is false, <> true for every non-NULL row */
		pc->x->constval = cbgpu_dict_lookup(dict, VARDATA_ANY(v), (int32) VARSIZE_ANY_EXHDR(v));
	}
	pending_consts = NIL;
	shim->cbestate->es_interconnect = cbgpu_shim_interconnect(shim_ctx, estate);
	if (has_motion && shim->cbestate->es_interconnect == NULL && getgpsegmentCount() > 1)
		SHIM_LOCAL_FAILURE(true, shim, "cbgpu: no GPU interconnect on segment %d", GpIdentity.segindex);
	shim->natts = desc->natts;
	shim->atttypes = (CbTypeId *) MemoryContextAllocZero(estate->es_query_cxt, sizeof(CbTypeId) * Max(desc->natts, 1));
	shim->attdscales = (int32 *) MemoryContextAllocZero(estate->es_query_cxt, sizeof(int32) * Max(desc->natts, 1));
	shim->attserial = (int *) MemoryContextAllocZero(estate->es_query_cxt, sizeof(int) * Max(desc->natts, 1));
	for (int a = 0; a < desc->natts; a++)
	{
		if (TupleDescAttr(desc, a)->atttypid == BYTEAOID && IsA(plan, Agg) && DO_AGGSPLIT_SERIALIZE(((Agg *) plan)->aggsplit) &&
			a < list_length(plan->targetlist) && IsA(((TargetEntry *) list_nth(plan->targetlist, a))->expr, Aggref))
		{
			/* a partial state on its way to a CPU Finalize stage: it leaves in the form the aggregate's serialisation function
			 * gives it (pg_aggregate.dat serialfn; cb_numeric_avg_serialize / cb_int8_avg_serialize make the same bytes from
			 * the device's (N, exact sum): tests/test_partial_state_serialize.py) */
			Aggref	   *ar = (Aggref *) ((TargetEntry *) list_nth(plan->targetlist, a))->expr;
			Oid			argtype = ar->aggargtypes != NIL ? linitial_oid(ar->aggargtypes) : InvalidOid;

			shim->attserial[a] = argtype == NUMERICOID ? 1 : argtype == INT8OID ? 2 : 0;
			shim->attdscales[a] = cplan->targetlist[a].expr->dscale;
			if (shim->attserial[a] != 0)
				continue;
		}
		if (!translate_type(TupleDescAttr(desc, a)->atttypid, TupleDescAttr(desc, a)->atttypmod, &shim->atttypes[a], &shim->attdscales[a]) &&
			TupleDescAttr(desc, a)->atttypid != NUMERICOID)
		{
			shim_release(shim);
			return NULL;
		}
	}
	/* device state goes away with the query context, error or not (utils/palloc.h MemoryContextRegisterResetCallback) */
	shim->reset_cb.func = shim_release;
	shim->reset_cb.arg = shim;
	MemoryContextRegisterResetCallback(estate->es_query_cxt, &shim->reset_cb);
	return shim;
}

/* route 2: try to take over the sub-tree rooted at ps; true when its ExecProcNode now points at the GPU path */
static bool
shim_take_over(PlanState *ps, EState *estate)
{
	CbgpuShim  *shim = shim_prepare(ps->plan, ps->ps_ResultTupleSlot ? ps->ps_ResultTupleSlot->tts_tupleDescriptor : NULL, estate);

	if (shim == NULL)
		return false;
	{
		ShimEntry  *e = (ShimEntry *) MemoryContextAllocZero(estate->es_query_cxt, sizeof(ShimEntry));

		e->ps = ps;
		e->shim = shim;
		e->next = shim_list;
		shim_list = e;
	}
	shim->saved_ExecProcNode = ps->ExecProcNodeReal;
	/* ExecProcNodeReal = ours, ExecProcNode = ExecProcNodeFirst -> ExecProcNodeGPDB (execProcnode.c:580-681): the reference's
	 * own wrapper keeps doing, around every call of ours, what it does for its own nodes - the QueryFinishPending early-out,
	 * the squelch check, the execprocnode DTrace probes, query_info_collect_hook(METRICS_PLAN_NODE_EXECUTING) and
	 * InstrStartNode / InstrStopNode, so EXPLAIN ANALYZE shows the replaced node's rows and time on the node it replaced */
	ExecSetExecProcNode(ps, cbgpu_ExecNode);
	return true;
}

/*
 * ExecReScan has no hook: a replaced node that were rescanned would have its (unused) CPU state reset by the node's own
 * ReScan function while the device state went on from where it was.  So a sub-tree is taken over only where the executor
 * cannot rescan it: nothing in it depends on a parameter (Plan.allParam empty: no correlated SubPlan / NestLoop parameter
 * will ever set chgParam on it, execAmi.c:77-120), and it does not sit on the inner side of a NestLoop or MergeJoin (rescanned
 * per outer row / restored to a mark).  SubPlans and InitPlans (PlanState.subPlan / initPlan) are not walked at all.
 */
static void
shim_walk(PlanState *ps, EState *estate, bool may_rescan)
{
	if (ps == NULL)
		return;
	if (!may_rescan && bms_is_empty(ps->plan->allParam) && shim_take_over(ps, estate))
		return;					/* the whole sub-tree is the GPU's: do not descend */
	shim_walk(outerPlanState(ps), estate, may_rescan);
	shim_walk(innerPlanState(ps), estate, may_rescan || IsA(ps, NestLoopState) || IsA(ps, MergeJoinState));
}

/* ------------------------------------------------------------------------------------------
 * route 1: the sub-tree as a CustomScan node (nodes/extensible.h:108-154, plannodes.h:1079).
 *
 * With cbgpu.route = 'customscan' the planner hook wraps every top-most translatable sub-tree of the finished plan in a
 * CustomScan: custom_plans keeps the original sub-tree (EXPLAIN shows it, and it is what BeginCustomScan translates),
 * custom_scan_tlist describes the tuple it returns, the node's own target list is plain INDEX_VAR references to that.
 * What this route has over the ExecProcNode swap (route 2): the executor knows the node - ReScanCustomScan gives
 * rescans (cb_ExecReScan resets the device sub-tree), EXPLAIN names it, and no PlanState pointer is overwritten.
 * Same device path underneath: shim_prepare / shim_next_tuple.
 * ------------------------------------------------------------------------------------------ */
typedef struct CbgpuScanState
{
	CustomScanState css;		/* must be first */
	CbgpuShim  *shim;
} CbgpuScanState;

static Node *cbgpu_create_scan_state(CustomScan *cscan);
static void cbgpu_begin_scan(CustomScanState *node, EState *estate, int eflags);
static TupleTableSlot *cbgpu_exec_scan(CustomScanState *node);
static void cbgpu_end_scan(CustomScanState *node);
static void cbgpu_rescan_scan(CustomScanState *node);

static const CustomScanMethods cbgpu_scan_methods = {"cbgpu", cbgpu_create_scan_state};
static const CustomExecMethods cbgpu_exec_methods = {
	.CustomName = "cbgpu",
	.BeginCustomScan = cbgpu_begin_scan,
	.ExecCustomScan = cbgpu_exec_scan,
	.EndCustomScan = cbgpu_end_scan,
	.ReScanCustomScan = cbgpu_rescan_scan,
};

static Node *
cbgpu_create_scan_state(CustomScan *cscan)
{
	CbgpuScanState *st = (CbgpuScanState *) newNode(sizeof(CbgpuScanState), T_CustomScanState);

	(void) cscan;
	st->css.methods = &cbgpu_exec_methods;
	return (Node *) st;
}

static void
cbgpu_begin_scan(CustomScanState *node, EState *estate, int eflags)
{
	CbgpuScanState *st = (CbgpuScanState *) node;
	CustomScan *cscan = (CustomScan *) node->ss.ps.plan;
	Plan	   *subtree = (Plan *) linitial(cscan->custom_plans);

	if (eflags & EXEC_FLAG_EXPLAIN_ONLY)
		return;
	/* the scan tuple is the sub-tree's result (custom_scan_tlist): ExecInitCustomScan built ss_ScanTupleSlot from it */
	st->shim = shim_prepare(subtree, node->ss.ss_ScanTupleSlot->tts_tupleDescriptor, estate);
	if (st->shim == NULL)
		ereport(ERROR, (errcode(ERRCODE_FEATURE_NOT_SUPPORTED),
						errmsg("cbgpu: the sub-tree wrapped at plan time cannot run on the device on segment %d", GpIdentity.segindex)));
}

static TupleTableSlot *
cbgpu_exec_scan(CustomScanState *node)
{
	CbgpuScanState *st = (CbgpuScanState *) node;
	TupleTableSlot *slot = shim_next_tuple(st->shim, node->ss.ss_ScanTupleSlot);

	if (TupIsNull(slot))
		return NULL;
	/* the node's own target list is one INDEX_VAR per scan column: project only if the planner put something else there */
	if (node->ss.ps.ps_ProjInfo)
	{
		node->ss.ps.ps_ExprContext->ecxt_scantuple = slot;
		return ExecProject(node->ss.ps.ps_ProjInfo);
	}
	return slot;
}

static void
cbgpu_end_scan(CustomScanState *node)
{
	CbgpuScanState *st = (CbgpuScanState *) node;

	if (st->shim)
		shim_release(st->shim);		/* idempotent: the query context's reset callback may run it again */
	st->shim = NULL;
}

static void
cbgpu_rescan_scan(CustomScanState *node)
{
	CbgpuScanState *st = (CbgpuScanState *) node;

	/* ExecReScan (execAmi.c:77) reaches us through ExecReScanCustomScan: start the device sub-tree over */
	if (st->shim && st->shim->cbps)
		cb_ExecReScan(st->shim->cbps);
}

/* can translate_plan take this sub-tree?  (plan-time check: no estate, nothing is loaded) */
static bool
subtree_translatable(Plan *plan)
{
	List	   *rels = NIL;
	CbPlan	   *c;

	if (plan == NULL || !bms_is_empty(plan->allParam) || !(IsA(plan, Agg) || IsA(plan, HashJoin) || IsA(plan, Motion)))
		return false;				/* bare scans are not worth a device round trip */
	pending_consts = NIL;
	c = translate_plan(plan, NULL, &rels);
	pending_consts = NIL;
	return c != NULL && resolve_plan(c);
}

static Plan *
wrap_subtrees(Plan *plan)
{
	if (plan == NULL)
		return NULL;
	if (subtree_translatable(plan))
	{
		CustomScan *cs = makeNode(CustomScan);
		ListCell   *lc;
		int			resno = 1;

		cs->scan.plan.startup_cost = plan->startup_cost;
		cs->scan.plan.total_cost = plan->total_cost;
		cs->scan.plan.plan_rows = plan->plan_rows;
		cs->scan.plan.plan_width = plan->plan_width;
		cs->scan.plan.flow = plan->flow;
		cs->scan.scanrelid = 0;			/* not a base relation scan: the tuple is described by custom_scan_tlist */
		cs->flags = 0;
		cs->custom_plans = list_make1(plan);
		cs->custom_scan_tlist = plan->targetlist;
		foreach(lc, plan->targetlist)
		{
			TargetEntry *te = (TargetEntry *) lfirst(lc);
			Var		   *v = makeVar(INDEX_VAR, resno, exprType((Node *) te->expr), exprTypmod((Node *) te->expr),
									exprCollation((Node *) te->expr), 0);

			cs->scan.plan.targetlist = lappend(cs->scan.plan.targetlist, makeTargetEntry((Expr *) v, resno, te->resname, te->resjunk));
			resno++;
		}
		cs->methods = &cbgpu_scan_methods;
		return &cs->scan.plan;
	}
	plan->lefttree = wrap_subtrees(plan->lefttree);
	/* the inner side of a NestLoop / MergeJoin is rescanned per outer row or restored to a mark: leave it to the CPU */
	if (!IsA(plan, NestLoop) && !IsA(plan, MergeJoin))
		plan->righttree = wrap_subtrees(plan->righttree);
	return plan;
}

static planner_hook_type prev_planner = NULL;
static char *cbgpu_route = NULL;		/* GUC cbgpu.route: "execprocnode" (default) or "customscan" */

static PlannedStmt *
cbgpu_planner(Query *parse, const char *query_string, int cursorOptions, ParamListInfo boundParams, OptimizerOptions *optimizer_options)
{
	PlannedStmt *stmt = prev_planner ? prev_planner(parse, query_string, cursorOptions, boundParams, optimizer_options)
		: standard_planner(parse, query_string, cursorOptions, boundParams, optimizer_options);

	if (stmt && stmt->commandType == CMD_SELECT && cbgpu_route && strcmp(cbgpu_route, "customscan") == 0)
		stmt->planTree = wrap_subtrees(stmt->planTree);
	return stmt;
}

/* dispatcher-side bookkeeping for the interconnect token: how many counted queries are between ExecutorStart and ExecutorEnd */
static int	qd_depth = 0;
static bool qd_failed = false;

static bool
qd_counts(CmdType operation, int eflags)
{
	return Gp_role == GP_ROLE_DISPATCH && operation == CMD_SELECT && !(eflags & EXEC_FLAG_EXPLAIN_ONLY);
}

static void
cbgpu_ExecutorEnd(QueryDesc *queryDesc)
{
	bool		counted = queryDesc->estate != NULL && qd_counts(queryDesc->operation, queryDesc->estate->es_top_eflags);

	if (prev_ExecutorEnd)
		prev_ExecutorEnd(queryDesc);
	else
		standard_ExecutorEnd(queryDesc);
	if (counted && qd_depth > 0)
		qd_depth--;
}

/* ereport(ERROR) unwinds past ExecutorEnd: the abort is where an interrupted query shows (access/xact.h XACT_EVENT_ABORT) */
static void
cbgpu_xact_callback(XactEvent event, void *arg)
{
	(void) arg;
	if (event == XACT_EVENT_ABORT || event == XACT_EVENT_PARALLEL_ABORT)
	{
		if (qd_depth > 0)
			qd_failed = true;
		qd_depth = 0;
	}
}

static void
cbgpu_ExecutorStart(QueryDesc *queryDesc, int eflags)
{
	/* on the dispatcher, before standard_ExecutorStart dispatches the plan: every QE must know the interconnect token.
	 * A query that died in error may have left a collective half done on some segment: the next one starts a new token */
	if (qd_counts(queryDesc->operation, eflags))
	{
		if (qd_depth == 0)
		{
			cbgpu_shim_qd_prepare(qd_failed);
			qd_failed = false;
		}
		qd_depth++;				/* nested queries (SPI in functions) keep the outer query's token */
	}
	if (prev_ExecutorStart)
		prev_ExecutorStart(queryDesc, eflags);
	else
		standard_ExecutorStart(queryDesc, eflags);
	/* route 2 swaps ExecProcNode pointers of the started plan; with route 1 the planner hook has wrapped the sub-trees already */
	if (!(eflags & EXEC_FLAG_EXPLAIN_ONLY) && !(cbgpu_route && strcmp(cbgpu_route, "customscan") == 0))
		shim_walk(queryDesc->planstate, queryDesc->estate, (eflags & (EXEC_FLAG_REWIND | EXEC_FLAG_BACKWARD | EXEC_FLAG_MARK)) != 0);
}

void
_PG_init(void)
{
	prev_ExecutorStart = ExecutorStart_hook;	/* executor/execMain.c:124 */
	ExecutorStart_hook = cbgpu_ExecutorStart;
	prev_ExecutorEnd = ExecutorEnd_hook;
	ExecutorEnd_hook = cbgpu_ExecutorEnd;
	RegisterXactCallback(cbgpu_xact_callback, NULL);
	cbgpu_shim_define_gucs();
	/* route 1: a CustomScan provider (RegisterCustomScanMethods lets the node be read back from a serialised plan on the
	 * QEs: nodes/readfuncs.c looks the methods up by name) and the planner hook that wraps sub-trees when asked to */
	RegisterCustomScanMethods(&cbgpu_scan_methods);
	DefineCustomStringVariable("cbgpu.route", "how GPU sub-trees enter the plan: execprocnode (default) or customscan", NULL,
							   &cbgpu_route, "execprocnode", PGC_USERSET, 0, NULL, NULL, NULL);
	prev_planner = planner_hook;
	planner_hook = cbgpu_planner;
}
