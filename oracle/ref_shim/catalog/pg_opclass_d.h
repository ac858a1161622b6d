/* stand-in for the generated catalog/pg_opclass_d.h (the reference build generates it; the files compiled against these stand-ins use
 * none of its constants) */
#ifndef STANDIN_PG_OPCLASS_D_H
#define STANDIN_PG_OPCLASS_D_H
#endif
