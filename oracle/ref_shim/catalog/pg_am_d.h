/* stand-in for the generated catalog/pg_am_d.h: the two access-method OIDs cdbhash.c names (catalog/pg_am.dat) */
#ifndef STANDIN_PG_AM_D_H
#define STANDIN_PG_AM_D_H
#define BTREE_AM_OID 403
#define HASH_AM_OID 405
#endif
