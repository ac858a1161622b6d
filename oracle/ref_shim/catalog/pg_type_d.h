/* stand-in for the genbki-generated catalog/pg_type_d.h: only the well-known constants the block format code needs */
#ifndef PG_TYPE_D_H
#define PG_TYPE_D_H
#define TYPALIGN_CHAR 'c'
#define TYPALIGN_SHORT 's'
#define TYPALIGN_INT 'i'
#define TYPALIGN_DOUBLE 'd'
#define TYPSTORAGE_PLAIN 'p'
#define TYPSTORAGE_EXTERNAL 'e'
#define TYPSTORAGE_EXTENDED 'x'
#define TYPSTORAGE_MAIN 'm'
#define BOOLOID 16
#define BYTEAOID 17
#define CHAROID 18
#define INT8OID 20
#define INT2OID 21
#define INT4OID 23
#define TEXTOID 25
#define OIDOID 26
#define FLOAT4OID 700
#define FLOAT8OID 701
#define BPCHAROID 1042
#define VARCHAROID 1043
#define DATEOID 1082
#define TIMEOID 1083
#define TIMESTAMPOID 1114
#define TIMESTAMPTZOID 1184
#define NUMERICOID 1700
#define RECORDOID 2249
#define INTERNALOID 2281
/* the EXPOSE_TO_CLIENT_CODE part of catalog/pg_type.h:280-300 that genbki copies here */
#define TYPTYPE_BASE 'b'
#define TYPTYPE_COMPOSITE 'c'
#define TYPTYPE_DOMAIN 'd'
#define TYPTYPE_ENUM 'e'
#define TYPTYPE_PSEUDO 'p'
#define TYPTYPE_RANGE 'r'
#endif
