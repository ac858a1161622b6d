/* stand-in for the generated utils/fmgrprotos.h (Gen_fmgrtab.pl): the builtin functions integration/cbgpu_shim.c calls */
#ifndef FMGRPROTOS_H
#define FMGRPROTOS_H
#include "fmgr.h"
extern Datum numeric_in(PG_FUNCTION_ARGS);
extern Datum numeric_out(PG_FUNCTION_ARGS);
extern Datum numeric_scale(PG_FUNCTION_ARGS);
extern Datum bpcharin(PG_FUNCTION_ARGS);
#endif
