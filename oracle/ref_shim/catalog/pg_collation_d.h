/* stand-in for the genbki-generated catalog/pg_collation_d.h: the well-known collation OIDs (catalog/pg_collation.dat:15-23) and
 * the EXPOSE_TO_CLIENT_CODE provider letters (catalog/pg_collation.h) that hashfunc.c / varchar.c name */
#ifndef PG_COLLATION_D_H
#define PG_COLLATION_D_H
#define DEFAULT_COLLATION_OID 100
#define C_COLLATION_OID 950
#define POSIX_COLLATION_OID 951
#define COLLPROVIDER_DEFAULT 'd'
#define COLLPROVIDER_ICU 'i'
#define COLLPROVIDER_LIBC 'c'
#endif
