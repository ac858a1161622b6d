/* stand-in for the generated catalog/pg_opfamily_d.h (the reference build generates it; the files compiled against these stand-ins use
 * none of its constants) */
#ifndef STANDIN_PG_OPFAMILY_D_H
#define STANDIN_PG_OPFAMILY_D_H
#endif
