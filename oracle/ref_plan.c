/*
 * ref_plan.c - exercising the f1 shim (integration/cbgpu_shim.c) on real reference Plan trees.  TEST INFRASTRUCTURE: only
 * tests/ loads the library this builds (oracle/_ref/libplan_ref.so); the product never does.
 *
 * The shim's translate_plan() takes the planner's output (nodes/plannodes.h Plan trees with primnodes.h expressions) and
 * produces the CbPlan the GPU executor runs.  No backend can be built here, so nothing ever handed it a real Plan - until
 * this file: it builds the plans the reference's planner emits for TPC-H Q1 / Q3 / Q5 (src/test/regress/expected/
 * aggregates.out:3313-3328 for the two-stage Q1 shape; rpt_tpch for Q3 / Q5) out of the reference's OWN node
 * constructors - nodes/makefuncs.c (makeVar, makeConst, makeTargetEntry, make_opclause, makeBoolExpr), nodes/list.c,
 * nodes/bitmapset.c, nodes/nodeFuncs.c (expression_tree_walker under pull_varattnos, optimizer/util/var.c), newNode from
 * nodes/nodes.c - with numeric literals made by the reference's numeric_in (utils/adt/numeric.c), all compiled where they
 * lie; operator and function OIDs come from the reference's catalog data (pg_operator.dat / pg_proc.dat through
 * gen_catalog_names.py), and the lsyscache lookups the shim makes (get_opname, get_func_name, get_func_rettype,
 * get_ordering_op_properties) are answered from the same data.  Then it runs the shim's translate_plan + resolve_plan on
 * them and hands the CbPlan and the scans' projected columns to the test, which executes it (oracle on the CPU, CUDA
 * executor on the GPU) against the reference's expected rows.
 *
 * What this pins: the shim reads the planner's structs the way the planner fills them - including what setrefs.c and
 * mark_partial_aggref leave behind (OUTER_VAR references, Aggref.args as TargetEntry lists, partial Aggrefs typed bytea,
 * Finalize Aggrefs over bytea Vars, Motion hash expressions over the child's target list).
 */
#include "../integration/cbgpu_shim.c"

#include "nodes/makefuncs.h"
#include "utils/date.h"

#include "ref_catalog_names.h"

#ifndef AGGKIND_NORMAL
#define AGGKIND_NORMAL 'n'			/* catalog/pg_aggregate.h:143 (genbki copies it into the generated pg_aggregate_d.h) */
#endif

/* ---- the catalog lookups the shim makes (utils/cache/lsyscache.c), from the reference's .dat files ---- */
static const RefOp *
op_by_oid(Oid oid)
{
	for (const RefOp *o = ref_ops; o->name; o++)
		if (o->oid == oid)
			return o;
	return NULL;
}

char *
get_opname(Oid opno)
{
	const RefOp *o = op_by_oid(opno);

	return o ? pstrdup(o->name) : NULL;
}

static const RefProc *
proc_by_oid(Oid oid)
{
	for (const RefProc *p = ref_procs; p->name; p++)
		if (p->oid == oid)
			return p;
	return NULL;
}

char *
get_func_name(Oid funcid)
{
	const RefProc *p = proc_by_oid(funcid);

	return p ? pstrdup(p->name) : NULL;
}

Oid
get_func_rettype(Oid funcid)
{
	const RefProc *p = proc_by_oid(funcid);

	return p ? p->rettype : InvalidOid;
}

bool
get_ordering_op_properties(Oid opno, Oid *opfamily, Oid *opcintype, int16 *strategy)
{
	const RefOp *o = op_by_oid(opno);

	if (!o || (strcmp(o->name, "<") != 0 && strcmp(o->name, ">") != 0))
		return false;
	*opfamily = InvalidOid;
	*opcintype = o->left;
	*strategy = o->name[0] == '<' ? BTLessStrategyNumber : BTGreaterStrategyNumber;
	return true;
}

/* ---- OIDs by name, as the parser's operator / function lookup would resolve them ---- */
static Oid
opno(const char *name, Oid left, Oid right)
{
	for (const RefOp *o = ref_ops; o->name; o++)
		if (strcmp(o->name, name) == 0 && o->left == left && o->right == right)
			return o->oid;
	elog(ERROR, "ref_plan: no operator %s(%u,%u)", name, left, right);
	return InvalidOid;
}

static Oid
aggfn(const char *name, Oid argtype, int nargs)
{
	for (const RefProc *p = ref_procs; p->name; p++)
		if (p->kind == 'a' && strcmp(p->name, name) == 0 && p->nargs == nargs && (nargs == 0 || p->argtype0 == argtype))
			return p->oid;
	/* count(*) is count() with no arguments; count(x) takes "any" */
	for (const RefProc *p = ref_procs; p->name; p++)
		if (p->kind == 'a' && strcmp(p->name, name) == 0 && p->nargs == nargs)
			return p->oid;
	elog(ERROR, "ref_plan: no aggregate %s(%u)", name, argtype);
	return InvalidOid;
}

/* ---- backend services the translate path uses (the contexts are ignored: the harness never frees) ---- */
void	   *MemoryContextAlloc(MemoryContext context, Size size) { return palloc(size); }
void	   *MemoryContextAllocZero(MemoryContext context, Size size) { return palloc0(size); }
void	   *MemoryContextAllocZeroAligned(MemoryContext context, Size size) { return palloc0(size); }
void		check_stack_depth(void) {}
ExecutorStart_hook_type ExecutorStart_hook = NULL;
ExecutorEnd_hook_type ExecutorEnd_hook = NULL;
planner_hook_type planner_hook = NULL;
GpId		GpIdentity = {0};
GpRoleValue Gp_role = GP_ROLE_UTILITY;

/* ---- the TPC-H tables as the reference's regression schema declares them (attribute numbers and types) ---- */
#define NUM152 (((15 << 16) | 2) + VARHDRSZ)	/* numeric(15,2) */
typedef struct Col
{
	const char *name;
	AttrNumber	attno;
	Oid			type;
	int32		typmod;
} Col;
static const Col lineitem_cols[] = {
	{"l_orderkey", 1, INT8OID, -1}, {"l_suppkey", 3, INT4OID, -1}, {"l_quantity", 5, NUMERICOID, NUM152},
	{"l_extendedprice", 6, NUMERICOID, NUM152}, {"l_discount", 7, NUMERICOID, NUM152}, {"l_tax", 8, NUMERICOID, NUM152},
	{"l_returnflag", 9, BPCHAROID, VARHDRSZ + 1}, {"l_linestatus", 10, BPCHAROID, VARHDRSZ + 1}, {"l_shipdate", 11, DATEOID, -1}, {0}};
static const Col orders_cols[] = {
	{"o_orderkey", 1, INT8OID, -1}, {"o_custkey", 2, INT4OID, -1}, {"o_orderdate", 5, DATEOID, -1}, {"o_shippriority", 8, INT4OID, -1}, {0}};
static const Col customer_cols[] = {
	{"c_custkey", 1, INT4OID, -1}, {"c_nationkey", 4, INT4OID, -1}, {"c_mktsegment", 7, BPCHAROID, VARHDRSZ + 10}, {0}};
static const Col supplier_cols[] = {{"s_suppkey", 1, INT4OID, -1}, {"s_nationkey", 4, INT4OID, -1}, {0}};
static const Col nation_cols[] = {{"n_nationkey", 1, INT4OID, -1}, {"n_name", 2, BPCHAROID, VARHDRSZ + 25}, {"n_regionkey", 3, INT4OID, -1}, {0}};
static const Col region_cols[] = {{"r_regionkey", 1, INT4OID, -1}, {"r_name", 2, BPCHAROID, VARHDRSZ + 25}, {0}};
/* range table order of the harness: 1 lineitem, 2 orders, 3 customer, 4 supplier, 5 nation, 6 region */
static const Col *const table_cols[] = {NULL, lineitem_cols, orders_cols, customer_cols, supplier_cols, nation_cols, region_cols};

static const Col *
col(Index rti, const char *name)
{
	for (const Col *c = table_cols[rti]; c->name; c++)
		if (strcmp(c->name, name) == 0)
			return c;
	elog(ERROR, "ref_plan: no column %s", name);
	return NULL;
}

static Var *
base_var(Index rti, const char *name)
{
	const Col  *c = col(rti, name);

	return makeVar(rti, c->attno, c->type, c->typmod, InvalidOid, 0);
}

/* a Var over the child's target entry `resno` (what set_upper_references leaves: varno OUTER_VAR / INNER_VAR) */
static Var *
child_var(Index varno, Plan *child, int resno)
{
	TargetEntry *te = (TargetEntry *) list_nth(child->targetlist, resno - 1);
	Node	   *e = (Node *) te->expr;
	Oid			type;
	int32		typmod = -1;

	if (IsA(e, Var))
	{
		type = ((Var *) e)->vartype;
		typmod = ((Var *) e)->vartypmod;
	}
	else if (IsA(e, Aggref))
		type = ((Aggref *) e)->aggtype;
	else if (IsA(e, OpExpr))
		type = ((OpExpr *) e)->opresulttype;
	else
	{
		elog(ERROR, "ref_plan: unexpected target entry");
		return NULL;
	}
	return makeVar(varno, resno, type, typmod, InvalidOid, 0);
}

static int	next_plan_id;

static void
set_tlist(Plan *p, List *exprs)
{
	ListCell   *lc;
	int			resno = 1;

	p->plan_node_id = ++next_plan_id;
	p->targetlist = NIL;
	foreach(lc, exprs)
	{
		p->targetlist = lappend(p->targetlist, makeTargetEntry((Expr *) lfirst(lc), resno, NULL, false));
		resno++;
	}
}

static Plan *
seqscan(Index rti, List *exprs, List *quals)
{
	SeqScan    *s = makeNode(SeqScan);

	s->scanrelid = rti;
	set_tlist(&s->plan, exprs);
	s->plan.qual = quals;
	return &s->plan;
}

static Const *
date_const(int32 days)
{
	return makeConst(DATEOID, -1, InvalidOid, 4, Int32GetDatum(days), false, true);
}

static Const *
numeric_const(const char *text)
{
	Datum		d = DirectFunctionCall3Coll(numeric_in, InvalidOid, CStringGetDatum(text), ObjectIdGetDatum(InvalidOid), Int32GetDatum(-1));

	return makeConst(NUMERICOID, -1, InvalidOid, -1, d, false, false);
}

static Const *
bpchar_const(const char *text)
{
	int			len = (int) strlen(text);
	struct varlena *v = (struct varlena *) palloc(VARHDRSZ + len);

	SET_VARSIZE(v, VARHDRSZ + len);
	memcpy(VARDATA(v), text, len);
	return makeConst(BPCHAROID, -1, InvalidOid, -1, PointerGetDatum(v), false, false);
}

static Expr *
op2(const char *name, Oid restype, Expr *l, Oid ltype, Expr *r, Oid rtype)
{
	return make_opclause(opno(name, ltype, rtype), restype, false, l, r, InvalidOid, InvalidOid);
}

/* one Aggref as the planner leaves it in an Agg's target list */
static Aggref *
aggref(const char *name, Oid argtype, Expr *arg, AggSplit split)
{
	Aggref	   *a = makeNode(Aggref);
	const RefProc *p;

	a->aggfnoid = aggfn(name, argtype, arg ? 1 : 0);
	p = proc_by_oid(a->aggfnoid);
	a->aggtype = p->rettype;
	a->aggargtypes = arg ? list_make1_oid(argtype) : NIL;
	a->args = arg ? list_make1(makeTargetEntry(arg, 1, NULL, false)) : NIL;
	a->aggstar = arg == NULL;
	a->aggkind = AGGKIND_NORMAL;
	a->aggsplit = split;
	if (DO_AGGSPLIT_SKIPFINAL(split))
	{
		/* mark_partial_aggref (optimizer/plan/planner.c): the partial result is the transition state; the numeric
		 * aggregates' and avg(int8)'s is `internal`, which travels serialised as bytea; count's is its int8 */
		if (strcmp(name, "count") != 0)
			a->aggtype = BYTEAOID;
	}
	return a;
}

/* sum(qty), sum(price), sum(price*(1-disc)), sum(price*(1-disc)*(1+tax)), avg(qty), avg(price), avg(disc), count(*) over
 * the scan's output columns 3..6 (rpt_tpch.source:346-371) */
static List *
q1_aggs(Plan *scan, AggSplit split)
{
	Expr	   *qty = (Expr *) child_var(OUTER_VAR, scan, 3);
	Expr	   *price = (Expr *) child_var(OUTER_VAR, scan, 4);
	Expr	   *disc = (Expr *) child_var(OUTER_VAR, scan, 5);
	Expr	   *tax = (Expr *) child_var(OUTER_VAR, scan, 6);
	Expr	   *one_minus = op2("-", NUMERICOID, (Expr *) numeric_const("1"), NUMERICOID, disc, NUMERICOID);
	Expr	   *rev = op2("*", NUMERICOID, price, NUMERICOID, one_minus, NUMERICOID);
	Expr	   *one_plus = op2("+", NUMERICOID, (Expr *) numeric_const("1"), NUMERICOID, tax, NUMERICOID);
	/* the same sub-expression again, as its own tree (the planner's target lists do not share nodes) */
	Expr	   *rev2 = op2("*", NUMERICOID, (Expr *) child_var(OUTER_VAR, scan, 4), NUMERICOID,
						   op2("-", NUMERICOID, (Expr *) numeric_const("1"), NUMERICOID, (Expr *) child_var(OUTER_VAR, scan, 5), NUMERICOID), NUMERICOID);
	Expr	   *chg = op2("*", NUMERICOID, rev2, NUMERICOID, one_plus, NUMERICOID);

	return list_make4(aggref("sum", NUMERICOID, qty, split), aggref("sum", NUMERICOID, price, split), aggref("sum", NUMERICOID, rev, split),
					  aggref("sum", NUMERICOID, chg, split));
}

static List *
q1_aggs_tail(Plan *scan, AggSplit split)
{
	Expr	   *qty = (Expr *) child_var(OUTER_VAR, scan, 3);
	Expr	   *price = (Expr *) child_var(OUTER_VAR, scan, 4);
	Expr	   *disc = (Expr *) child_var(OUTER_VAR, scan, 5);

	return list_make4(aggref("avg", NUMERICOID, qty, split), aggref("avg", NUMERICOID, price, split), aggref("avg", NUMERICOID, disc, split),
					  aggref("count", InvalidOid, NULL, split));
}

static Agg *
make_agg(Plan *child, AggStrategy strategy, AggSplit split, int ngroup, List *tlist)
{
	Agg		   *a = makeNode(Agg);

	a->aggstrategy = strategy;
	a->aggsplit = split;
	a->numCols = ngroup;
	a->grpColIdx = (AttrNumber *) palloc0(sizeof(AttrNumber) * Max(ngroup, 1));
	for (int i = 0; i < ngroup; i++)
		a->grpColIdx[i] = i + 1;
	a->numGroups = 6;
	a->plan.lefttree = child;
	set_tlist(&a->plan, tlist);
	return a;
}

static Motion *
make_motion(Plan *child, MotionType type, List *hashExprs, int nsegs)
{
	Motion	   *m = makeNode(Motion);
	List	   *tl = NIL;

	m->motionType = type;
	m->motionID = ++next_plan_id;
	m->hashExprs = hashExprs;
	m->numHashSegments = type == MOTIONTYPE_HASH ? nsegs : 0;
	m->plan.lefttree = child;
	for (int i = 1; i <= list_length(child->targetlist); i++)
		tl = lappend(tl, child_var(OUTER_VAR, child, i));
	set_tlist(&m->plan, tl);
	return m;
}

static Plan *
build_q1(int nsegs, int32 cutoff)
{
	Plan	   *scan = seqscan(1, lappend(list_make5(base_var(1, "l_returnflag"), base_var(1, "l_linestatus"), base_var(1, "l_quantity"),
													  base_var(1, "l_extendedprice"), base_var(1, "l_discount")), base_var(1, "l_tax")),
							   list_make1(op2("<=", BOOLOID, (Expr *) base_var(1, "l_shipdate"), DATEOID, (Expr *) date_const(cutoff), DATEOID)));
	List	   *keys = list_make2(child_var(OUTER_VAR, scan, 1), child_var(OUTER_VAR, scan, 2));

	if (nsegs == 1)
		return &make_agg(scan, AGG_HASHED, AGGSPLIT_SIMPLE, 2,
						 list_concat(list_concat(keys, q1_aggs(scan, AGGSPLIT_SIMPLE)), q1_aggs_tail(scan, AGGSPLIT_SIMPLE)))->plan;
	{
		/* Gather Motion <- Finalize HashAggregate <- Redistribute Motion <- Partial HashAggregate <- Seq Scan */
		Agg		   *partial = make_agg(scan, AGG_HASHED, AGGSPLIT_INITIAL_SERIAL, 2,
									   list_concat(list_concat(keys, q1_aggs(scan, AGGSPLIT_INITIAL_SERIAL)), q1_aggs_tail(scan, AGGSPLIT_INITIAL_SERIAL)));
		Motion	   *redist;
		List	   *ftl;
		Agg		   *final;
		static const char *const names[8] = {"sum", "sum", "sum", "sum", "avg", "avg", "avg", "count"};

		partial->streaming = true;
		redist = make_motion(&partial->plan, MOTIONTYPE_HASH,
							 list_make2(child_var(OUTER_VAR, &partial->plan, 1), child_var(OUTER_VAR, &partial->plan, 2)), nsegs);
		ftl = list_make2(child_var(OUTER_VAR, &redist->plan, 1), child_var(OUTER_VAR, &redist->plan, 2));
		for (int i = 0; i < 8; i++)
		{
			/* the Finalize Aggref: its argument is the Var carrying the partial state (bytea, or count's int8) */
			/* ... and it is otherwise the original Aggref: count(*) keeps aggfnoid 2803 and aggstar (the planner copies
			 * the Aggref and only swaps its arguments: make_partial_grouping_target / convert_combining_aggrefs) */
			Aggref	   *a = aggref(names[i], i == 7 ? InvalidOid : NUMERICOID, i == 7 ? NULL : (Expr *) child_var(OUTER_VAR, &redist->plan, 3 + i),
								   AGGSPLIT_FINAL_DESERIAL);

			if (i == 7)
				a->args = list_make1(makeTargetEntry((Expr *) child_var(OUTER_VAR, &redist->plan, 3 + i), 1, NULL, false));
			ftl = lappend(ftl, a);
		}
		final = make_agg(&redist->plan, AGG_HASHED, AGGSPLIT_FINAL_DESERIAL, 2, ftl);
		return &make_motion(&final->plan, MOTIONTYPE_GATHER, NIL, nsegs)->plan;
	}
}

static Hash *
make_hash(Plan *child, List *hashkeys)
{
	Hash	   *h = makeNode(Hash);
	List	   *tl = NIL;

	h->plan.lefttree = child;
	h->hashkeys = hashkeys;
	for (int i = 1; i <= list_length(child->targetlist); i++)
		tl = lappend(tl, child_var(OUTER_VAR, child, i));
	set_tlist(&h->plan, tl);
	return h;
}

static HashJoin *
make_hashjoin(JoinType jt, Plan *outer, Hash *inner, List *hashkeys, List *tlist)
{
	HashJoin   *j = makeNode(HashJoin);

	j->join.jointype = jt;
	j->join.plan.lefttree = outer;
	j->join.plan.righttree = &inner->plan;
	j->hashkeys = hashkeys;
	set_tlist(&j->join.plan, tlist);
	return j;
}

/* HashAggregate(l_orderkey, o_orderdate, o_shippriority; sum(l_extendedprice * (1 - l_discount)))
 *   <- Hash Join (l_orderkey = o_orderkey) <- Seq Scan lineitem (l_shipdate > d)
 *        <- Hash <- Hash Join (o_custkey = c_custkey) <- Seq Scan orders (o_orderdate < d)
 *                     <- Hash <- Seq Scan customer (c_mktsegment = 'MACHINERY')
 * (rpt_tpch.source:458-480; the Sort / Limit above stay with the CPU executor: the shim takes the sub-tree under them) */
static Plan *
build_q3(const char *segment, int32 cutoff)
{
	Plan	   *cust = seqscan(3, list_make1(base_var(3, "c_custkey")),
							   list_make1(op2("=", BOOLOID, (Expr *) base_var(3, "c_mktsegment"), BPCHAROID, (Expr *) bpchar_const(segment), BPCHAROID)));
	Plan	   *orders = seqscan(2, list_make4(base_var(2, "o_orderkey"), base_var(2, "o_custkey"), base_var(2, "o_orderdate"), base_var(2, "o_shippriority")),
								 list_make1(op2("<", BOOLOID, (Expr *) base_var(2, "o_orderdate"), DATEOID, (Expr *) date_const(cutoff), DATEOID)));
	Hash	   *hc = make_hash(cust, list_make1(child_var(OUTER_VAR, cust, 1)));
	HashJoin   *j1 = make_hashjoin(JOIN_INNER, orders, hc, list_make1(child_var(OUTER_VAR, orders, 2)),
								   list_make3(child_var(OUTER_VAR, orders, 1), child_var(OUTER_VAR, orders, 3), child_var(OUTER_VAR, orders, 4)));
	Plan	   *li = seqscan(1, list_make3(base_var(1, "l_orderkey"), base_var(1, "l_extendedprice"), base_var(1, "l_discount")),
							 list_make1(op2(">", BOOLOID, (Expr *) base_var(1, "l_shipdate"), DATEOID, (Expr *) date_const(cutoff), DATEOID)));
	Hash	   *ho = make_hash(&j1->join.plan, list_make1(child_var(OUTER_VAR, &j1->join.plan, 1)));
	HashJoin   *j2 = make_hashjoin(JOIN_INNER, li, ho, list_make1(child_var(OUTER_VAR, li, 1)),
								   list_make5(child_var(OUTER_VAR, li, 1), child_var(INNER_VAR, &ho->plan, 2), child_var(INNER_VAR, &ho->plan, 3),
											  child_var(OUTER_VAR, li, 2), child_var(OUTER_VAR, li, 3)));
	Plan	   *j = &j2->join.plan;
	Expr	   *rev = op2("*", NUMERICOID, (Expr *) child_var(OUTER_VAR, j, 4), NUMERICOID,
						  op2("-", NUMERICOID, (Expr *) numeric_const("1"), NUMERICOID, (Expr *) child_var(OUTER_VAR, j, 5), NUMERICOID), NUMERICOID);
	Agg		   *a = make_agg(j, AGG_HASHED, AGGSPLIT_SIMPLE, 3,
							 list_make4(child_var(OUTER_VAR, j, 1), child_var(OUTER_VAR, j, 2), child_var(OUTER_VAR, j, 3),
										aggref("sum", NUMERICOID, rev, AGGSPLIT_SIMPLE)));

	a->numGroups = 1000000;
	/* grouping columns 1..3 of the Agg's child; the Agg's own target list is (l_orderkey, o_orderdate, o_shippriority, revenue) */
	return &a->plan;
}

/* HashAggregate(n_name; sum(l_extendedprice * (1 - l_discount))) over the six-table join of rpt_tpch.source:512-535:
 *   lineitem ⨝ orders(date range) ⨝ customer ⨝ supplier (s_suppkey = l_suppkey AND s_nationkey = c_nationkey) ⨝ nation ⨝ region('AMERICA') */
static Plan *
build_q5(const char *region, int32 date_lo, int32 date_hi)
{
	Plan	   *reg = seqscan(6, list_make1(base_var(6, "r_regionkey")),
							  list_make1(op2("=", BOOLOID, (Expr *) base_var(6, "r_name"), BPCHAROID, (Expr *) bpchar_const(region), BPCHAROID)));
	Plan	   *nat = seqscan(5, list_make3(base_var(5, "n_nationkey"), base_var(5, "n_name"), base_var(5, "n_regionkey")), NIL);
	Hash	   *hr = make_hash(reg, list_make1(child_var(OUTER_VAR, reg, 1)));
	HashJoin   *jn = make_hashjoin(JOIN_INNER, nat, hr, list_make1(child_var(OUTER_VAR, nat, 3)),
								   list_make2(child_var(OUTER_VAR, nat, 1), child_var(OUTER_VAR, nat, 2)));
	Plan	   *sup = seqscan(4, list_make2(base_var(4, "s_suppkey"), base_var(4, "s_nationkey")), NIL);
	Hash	   *hn = make_hash(&jn->join.plan, list_make1(child_var(OUTER_VAR, &jn->join.plan, 1)));
	/* supplier ⨝ nation: (s_suppkey, s_nationkey, n_name) */
	HashJoin   *js = make_hashjoin(JOIN_INNER, sup, hn, list_make1(child_var(OUTER_VAR, sup, 2)),
								   list_make3(child_var(OUTER_VAR, sup, 1), child_var(OUTER_VAR, sup, 2), child_var(INNER_VAR, &hn->plan, 2)));
	Plan	   *cust = seqscan(3, list_make2(base_var(3, "c_custkey"), base_var(3, "c_nationkey")), NIL);
	Plan	   *ord = seqscan(2, list_make2(base_var(2, "o_orderkey"), base_var(2, "o_custkey")),
							  list_make2(op2(">=", BOOLOID, (Expr *) base_var(2, "o_orderdate"), DATEOID, (Expr *) date_const(date_lo), DATEOID),
										 op2("<", BOOLOID, (Expr *) base_var(2, "o_orderdate"), DATEOID, (Expr *) date_const(date_hi), DATEOID)));
	Hash	   *hcu = make_hash(cust, list_make1(child_var(OUTER_VAR, cust, 1)));
	/* orders ⨝ customer: (o_orderkey, c_nationkey) */
	HashJoin   *jo = make_hashjoin(JOIN_INNER, ord, hcu, list_make1(child_var(OUTER_VAR, ord, 2)),
								   list_make2(child_var(OUTER_VAR, ord, 1), child_var(INNER_VAR, &hcu->plan, 2)));
	Plan	   *li = seqscan(1, list_make4(base_var(1, "l_orderkey"), base_var(1, "l_suppkey"), base_var(1, "l_extendedprice"), base_var(1, "l_discount")), NIL);
	Hash	   *ho = make_hash(&jo->join.plan, list_make1(child_var(OUTER_VAR, &jo->join.plan, 1)));
	/* lineitem ⨝ (orders ⨝ customer): (l_suppkey, c_nationkey, l_extendedprice, l_discount) */
	HashJoin   *jl = make_hashjoin(JOIN_INNER, li, ho, list_make1(child_var(OUTER_VAR, li, 1)),
								   list_make4(child_var(OUTER_VAR, li, 2), child_var(INNER_VAR, &ho->plan, 2), child_var(OUTER_VAR, li, 3),
											  child_var(OUTER_VAR, li, 4)));
	Hash	   *hs = make_hash(&js->join.plan, list_make2(child_var(OUTER_VAR, &js->join.plan, 1), child_var(OUTER_VAR, &js->join.plan, 2)));
	/* ... ⨝ supplier on (l_suppkey, c_nationkey) = (s_suppkey, s_nationkey): (n_name, l_extendedprice, l_discount) */
	HashJoin   *jt = make_hashjoin(JOIN_INNER, &jl->join.plan, hs,
								   list_make2(child_var(OUTER_VAR, &jl->join.plan, 1), child_var(OUTER_VAR, &jl->join.plan, 2)),
								   list_make3(child_var(INNER_VAR, &hs->plan, 3), child_var(OUTER_VAR, &jl->join.plan, 3),
											  child_var(OUTER_VAR, &jl->join.plan, 4)));
	Plan	   *j = &jt->join.plan;
	Expr	   *rev = op2("*", NUMERICOID, (Expr *) child_var(OUTER_VAR, j, 2), NUMERICOID,
						  op2("-", NUMERICOID, (Expr *) numeric_const("1"), NUMERICOID, (Expr *) child_var(OUTER_VAR, j, 3), NUMERICOID), NUMERICOID);
	Agg		   *a = make_agg(j, AGG_HASHED, AGGSPLIT_SIMPLE, 1, list_make2(child_var(OUTER_VAR, j, 1), aggref("sum", NUMERICOID, rev, AGGSPLIT_SIMPLE)));

	a->numGroups = 25;
	return &a->plan;
}

/* ------------------------------------------------------------------------------------------
 * the test's entry points
 * ------------------------------------------------------------------------------------------ */
extern jmp_buf *ref_exec_jmp(void);		/* ref_exec.c: ereport(ERROR) and stubbed backend functions land here */
extern const char *ref_exec_last_error(void);

static List *last_rels;
static List *last_pending;

/* which: "q1" (a = segments, b = cutoff date), "q3" (text = market segment, b = cutoff), "q5" (text = region, b / c = date
 * range).  Returns the translated CbPlan (palloc'ed, lives until the process ends) or NULL: ref_plan_error() says why. */
CbPlan *
ref_plan_translate(const char *which, const char *text, int a, int b, int c)
{
	Plan	   *p;
	CbPlan	   *out;

	if (setjmp(*ref_exec_jmp()) != 0)
		return NULL;
	next_plan_id = 0;
	if (strcmp(which, "q1") == 0)
		p = build_q1(a, b);
	else if (strcmp(which, "q3") == 0)
		p = build_q3(text, b);
	else if (strcmp(which, "q5") == 0)
		p = build_q5(text, b, c);
	else
		return NULL;
	last_rels = NIL;
	pending_consts = NIL;
	out = translate_plan(p, NULL, &last_rels);
	if (out && !resolve_plan(out))
		out = NULL;
	last_pending = pending_consts;
	return out;
}

/* Route 1: what the planner hook does to the same plans - wrap_subtrees() must put ONE CustomScan on top (every harness plan
 * is translatable as a whole), keep the original sub-tree in custom_plans, describe the tuple with custom_scan_tlist and give the
 * node a target list of INDEX_VAR references of the right types.  Returns the number of columns, or a negative code. */
int
ref_plan_wrap(const char *which, const char *text, int a, int b, int c)
{
	Plan	   *p,
			   *w;
	CustomScan *cs;
	ListCell   *lc,
			   *lo;
	int			n = 0;

	if (setjmp(*ref_exec_jmp()) != 0)
		return -100;
	next_plan_id = 0;
	if (strcmp(which, "q1") == 0)
		p = build_q1(a, b);
	else if (strcmp(which, "q3") == 0)
		p = build_q3(text, b);
	else if (strcmp(which, "q5") == 0)
		p = build_q5(text, b, c);
	else
		return -1;
	w = wrap_subtrees(p);
	if (w == NULL || !IsA(w, CustomScan))
		return -2;
	cs = (CustomScan *) w;
	if (list_length(cs->custom_plans) != 1 || linitial(cs->custom_plans) != (void *) p || cs->custom_scan_tlist != p->targetlist ||
		cs->methods != &cbgpu_scan_methods || cs->scan.scanrelid != 0)
		return -3;
	if (list_length(cs->scan.plan.targetlist) != list_length(p->targetlist))
		return -4;
	forboth(lc, cs->scan.plan.targetlist, lo, p->targetlist)
	{
		TargetEntry *te = (TargetEntry *) lfirst(lc);
		Var		   *v = (Var *) te->expr;

		n++;
		if (!IsA(v, Var) || v->varno != INDEX_VAR || v->varattno != n || te->resno != n ||
			v->vartype != exprType((Node *) ((TargetEntry *) lfirst(lo))->expr))
			return -5;
	}
	/* and what BeginCustomScan would translate is still the plan the other route takes */
	last_rels = NIL;
	pending_consts = NIL;
	if (translate_plan((Plan *) linitial(cs->custom_plans), NULL, &last_rels) == NULL)
		return -6;
	pending_consts = NIL;
	return n;
}

/* the i-th scan of the translated tree (= range-table index i + 1 of the CbPlan): the reference range-table index it reads
 * and the attribute numbers it projects, ascending = the device relation's column order.  Returns the column count. */
int
ref_plan_scan(int i, int *rti, int *attnos, int cap)
{
	ShimScan   *sc;
	ListCell   *lc;
	int			n = 0;

	if (i < 0 || i >= list_length(last_rels))
		return -1;
	sc = (ShimScan *) list_nth(last_rels, i);
	*rti = (int) sc->scanrelid;
	foreach(lc, sc->attnos)
	{
		if (n < cap)
			attnos[n] = lfirst_int(lc);
		n++;
	}
	return n;
}

/* string literals compared with dictionary columns: the loader turns them into codes once the relation's dictionary exists
 * (shim_take_over does it through cbgpu_shim_dict); the test plays the loader.  Returns the literal's length, -1 past the end. */
int
ref_plan_pending(int i, int *rti, int *attno, char *text, int cap)
{
	ShimPendingConst *pc;
	struct varlena *v;
	int			len;

	if (i < 0 || i >= list_length(last_pending))
		return -1;
	pc = (ShimPendingConst *) list_nth(last_pending, i);
	*rti = (int) pc->scanrelid;
	*attno = pc->attno;
	v = (struct varlena *) DatumGetPointer(pc->c->constvalue);
	len = (int) VARSIZE_ANY_EXHDR(v);
	if (len < cap)
	{
		memcpy(text, VARDATA_ANY(v), len);
		text[len] = 0;
	}
	return len;
}

void
ref_plan_resolve_pending(int i, int64 code)
{
	if (i >= 0 && i < list_length(last_pending))
		((ShimPendingConst *) list_nth(last_pending, i))->x->constval = code;
}

const char *
ref_plan_error(void)
{
	return ref_exec_last_error();
}
