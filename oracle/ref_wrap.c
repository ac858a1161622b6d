/*
 * oracle/ref_wrap.c - exports the reference's static-inline hash helpers so tests can call them.
 * Compiled (by oracle/Makefile, only where /root/reference exists) together with the reference's
 * own src/common/hashfn.c into oracle/_ref/libpg_hashfn.so.  No reference source is copied: the
 * functions come from the reference headers at build time.  TEST INFRASTRUCTURE ONLY.
 */
#include "postgres.h"
#include "common/hashfn.h"

uint32		ref_murmurhash32(uint32 x);
uint32		ref_hash_combine32(uint32 a, uint32 b);

uint32
ref_murmurhash32(uint32 x)
{
	return murmurhash32(x);		/* src/include/common/hashfn.h:93 */
}

uint32
ref_hash_combine32(uint32 a, uint32 b)
{
	return hash_combine(a, b);	/* src/include/common/hashfn.h:80 */
}
