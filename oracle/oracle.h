/*
 * oracle/oracle.h - API of the CPU oracle: a row-at-a-time restatement of the reference's
 * scan -> hash join -> hash aggregate (+ Redistribute Motion) executor path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md).  It consumes the same plan structs
 * (include/cb_plan.h) and the same column encodings as the product, so a parity test feeds both
 * the identical plan and data.
 */
#ifndef ORACLE_H
#define ORACLE_H

#include <stdint.h>
#include "../include/cb_plan.h"

#ifdef __cplusplus
extern "C" {
#endif

/* a host relation: fixed-width column arrays (the decoded form of an AOCS table) */
typedef struct OraRel
{
	int64_t		nrows;
	int32_t		ncols;
	int32_t	   *types;		/* CbTypeId per column                                          */
	int32_t	   *dscales;	/* numeric display scale per column                             */
	void	  **data;		/* column arrays, element width = cb_type_width(type)           */
	uint8_t	  **nulls;		/* per column: NULL, or one byte per row (1 = NULL)             */
	uint8_t	   *visimap;	/* NULL, or one BIT per row (1 = visible), appendonly_visimap.c */
	uint32_t  **dict_hash;	/* per column: for CB_DICT8/32, hashbpchar() of each code's text */
} OraRel;

/* one "segment" = one range table (relations indexed by scanrelid-1) */
typedef struct OraSegment
{
	int32_t		nrels;
	OraRel	  **rels;
} OraSegment;

typedef struct OraResult OraResult;

/*
 * Execute `plan` on a simulated cluster of `nsegs` segments (one range table each).  Motion nodes
 * route rows between the segments exactly as nodeMotion.c / cdbhash.c would.  The rows reaching the
 * top of the plan on every segment are concatenated (segment order) into the result.
 * nthreads > 1 runs segments concurrently below each Motion (one thread per segment at a time).
 * Returns NULL on error; ora_last_error() has the message.
 */
OraResult  *ora_execute(const CbPlan *plan, OraSegment *segs, int32_t nsegs, int32_t nthreads);
const char *ora_last_error(void);

int64_t		ora_result_nrows(const OraResult *r);
int32_t		ora_result_ncols(const OraResult *r);
int32_t		ora_result_type(const OraResult *r, int32_t col);
int32_t		ora_result_segment(const OraResult *r, int64_t row);	/* segment that emitted the row */
/* value accessors; numeric results are exact decimal text formatted as numeric_out() prints them */
int			ora_result_isnull(const OraResult *r, int64_t row, int32_t col);
int64_t		ora_result_int64(const OraResult *r, int64_t row, int32_t col);
double		ora_result_float8(const OraResult *r, int64_t row, int32_t col);
const char *ora_result_text(const OraResult *r, int64_t row, int32_t col);
/* partial aggregate states (AGGSPLIT_INITIAL_SERIAL outputs): N and the 128-bit sum */
int64_t		ora_result_state_n(const OraResult *r, int64_t row, int32_t col);
void		ora_result_state_sum(const OraResult *r, int64_t row, int32_t col, int64_t *lo, int64_t *hi);
void		ora_result_free(OraResult *r);

/* standalone pieces, for operator-level parity tests */
uint32_t	ora_hash_datum(int32_t type, int64_t value_bits);				/* hashint4/8, hashfloat8, bpchar(1) */
uint32_t	ora_hashbpchar_text(const char *s, int32_t len);
int32_t		ora_cdbhash_segment(const int32_t *types, const int64_t *values, const uint8_t *isnull,
								 int32_t nkeys, int32_t numsegs);
/* numeric finalisation (numeric_sum / numeric_avg text) from an exact (sum, N) pair */
void		ora_numeric_sum_text(int64_t lo, int64_t hi, int32_t dscale, char *out, int32_t outlen);
void		ora_numeric_avg_text(int64_t lo, int64_t hi, int32_t dscale, int64_t n, char *out, int32_t outlen);

#ifdef __cplusplus
}
#endif
#endif
