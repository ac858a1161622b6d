/* stand-in for the generated catalog/pg_proc_d.h (the reference build generates it; the files compiled or type-checked against
 * these stand-ins use none of its constants) */
#ifndef STANDIN_PG_PROC_D_H
#define STANDIN_PG_PROC_D_H
#endif
