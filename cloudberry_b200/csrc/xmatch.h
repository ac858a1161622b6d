/*
 * xmatch.h - host-side reconstruction of expression trees from a pipeline's postfix program, used
 * by the pattern matchers that route a descriptor to a hand-fused kernel (the role the reference's
 * ExecReadyInterpretedExpr / EEOP fast-path selection plays for common expression shapes,
 * backend/executor/execExprInterp.c:260-330).
 */
#pragma once
#include "common.cuh"

#define XM_MAXNODES 320

struct XNode
{
	int			code;			/* CbpOpCode */
	int			a;
	int64_t		imm;
	int			l, r;
};

struct XSection
{
	int			kind;			/* 0 = FILTER (root node in `node`), 1 = PROBE (`probe`, keys in `keys`) */
	int			node;
	int			probe;
	int			keys[CBP_MAX_KEYS];
};

struct XProg
{
	XNode		nodes[XM_MAXNODES];
	int			n;
	XSection	sections[CBP_MAX_OPS];
	int			nsections;
	int			stack[CBP_STACK];	/* node ids left on the stack at END: keys then args / outputs    */
	int			depth;
};

static inline bool
xm_decompile(const CbPipeline *p, XProg *x)
{
	x->n = 0;
	x->nsections = 0;
	x->depth = 0;
	for (int i = 0; i < p->nops; i++)
	{
		const CbpOp *op = &p->ops[i];

		switch (op->code)
		{
			case CBP_LOAD:
			case CBP_CONST:
				if (x->n >= XM_MAXNODES)
					return false;
				x->nodes[x->n] = {op->code, op->a, op->imm, -1, -1};
				x->stack[x->depth++] = x->n++;
				break;
			case CBP_DUP:
				x->stack[x->depth] = x->stack[op->a];
				x->depth++;
				break;
			case CBP_POP:
				x->depth--;
				break;
			case CBP_F8ORD:
				return false;		/* float arithmetic stays with the generic kernel */
			case CBP_NOT:
			case CBP_I2F:
				if (x->n >= XM_MAXNODES)
					return false;
				x->nodes[x->n] = {op->code, op->a, 0, x->stack[x->depth - 1], -1};
				x->stack[x->depth - 1] = x->n++;
				break;
			case CBP_FILTER:
				x->depth--;
				x->sections[x->nsections].kind = 0;
				x->sections[x->nsections].node = x->stack[x->depth];
				x->nsections++;
				break;
			case CBP_PROBE:
				{
					int			nk = p->probes[op->a].nkeys;

					x->depth -= nk;
					x->sections[x->nsections].kind = 1;
					x->sections[x->nsections].probe = op->a;
					for (int k = 0; k < nk; k++)
						x->sections[x->nsections].keys[k] = x->stack[x->depth + k];
					x->nsections++;
					break;
				}
			case CBP_END:
				break;
			default:			/* binary operators */
				if (x->n >= XM_MAXNODES)
					return false;
				x->nodes[x->n] = {op->code, op->a, 0, x->stack[x->depth - 2], x->stack[x->depth - 1]};
				x->depth--;
				x->stack[x->depth - 1] = x->n++;
				break;
		}
	}
	return true;
}

static inline bool
xm_is_load(const XProg *x, int node, int *col)
{
	if (node < 0 || x->nodes[node].code != CBP_LOAD)
		return false;
	*col = x->nodes[node].a;
	return true;
}

static inline bool
xm_is_const(const XProg *x, int node, int64_t *v)
{
	if (node < 0 || x->nodes[node].code != CBP_CONST)
		return false;
	*v = x->nodes[node].imm;
	return true;
}

/* MUL(LOAD b, SUB(CONST k, LOAD c))  -  e.g. l_extendedprice * (1 - l_discount) */
static inline bool
xm_is_rev(const XProg *x, int node, int *b, int64_t *k, int *c)
{
	if (node < 0 || x->nodes[node].code != CBP_MUL)
		return false;
	int			s = x->nodes[node].r;

	if (!xm_is_load(x, x->nodes[node].l, b) || s < 0 || x->nodes[s].code != CBP_SUB)
		return false;
	return xm_is_const(x, x->nodes[s].l, k) && xm_is_load(x, x->nodes[s].r, c);
}

/* MUL(rev, ADD(CONST k2, LOAD d))  -  e.g. ... * (1 + l_tax) */
static inline bool
xm_is_chg(const XProg *x, int node, int *b, int64_t *k, int *c, int64_t *k2, int *d)
{
	if (node < 0 || x->nodes[node].code != CBP_MUL)
		return false;
	int			s = x->nodes[node].r;

	if (!xm_is_rev(x, x->nodes[node].l, b, k, c) || s < 0 || x->nodes[s].code != CBP_ADD)
		return false;
	return xm_is_const(x, x->nodes[s].l, k2) && xm_is_load(x, x->nodes[s].r, d);
}

/* CMP(LOAD col, CONST v) */
static inline bool
xm_is_cmp_const(const XProg *x, int node, int *code, int *col, int64_t *v)
{
	if (node < 0)
		return false;
	int			c = x->nodes[node].code;

	if (c < CBP_EQ || c > CBP_GE)
		return false;
	*code = c;
	return xm_is_load(x, x->nodes[node].l, col) && xm_is_const(x, x->nodes[node].r, v);
}
