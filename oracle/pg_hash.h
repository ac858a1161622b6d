/*
 * oracle/pg_hash.h - CPU restatement of the reference's hash functions.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or executed by the product
 * path (cloudberry_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, as the checker or the CPU baseline.
 *
 * Each function follows the reference function cited beside it.  Pinned against (a) the known
 * answers SURVEY.md 8c lists (computed from the reference's own hashfn.o) and (b), when
 * /root/reference is present, the reference's src/common/hashfn.c compiled as
 * oracle/_ref/libpg_hashfn.so (tests/test_oracle_hash.py).
 */
#ifndef ORACLE_PG_HASH_H
#define ORACLE_PG_HASH_H

#include <stdint.h>
#include <string.h>

#define ORA_ROT(x, k) (((x) << (k)) | ((x) >> (32 - (k))))

/* src/common/hashfn.c:81-89 mix() */
#define ORA_MIX(a, b, c) \
	do { \
		a -= c; a ^= ORA_ROT(c, 4); c += b; \
		b -= a; b ^= ORA_ROT(a, 6); a += c; \
		c -= b; c ^= ORA_ROT(b, 8); b += a; \
		a -= c; a ^= ORA_ROT(c, 16); c += b; \
		b -= a; b ^= ORA_ROT(a, 19); a += c; \
		c -= b; c ^= ORA_ROT(b, 4); b += a; \
	} while (0)

/* src/common/hashfn.c:133-142 final() */
#define ORA_FINAL(a, b, c) \
	do { \
		c ^= b; c -= ORA_ROT(b, 14); \
		a ^= c; a -= ORA_ROT(c, 11); \
		b ^= a; b -= ORA_ROT(a, 25); \
		c ^= b; c -= ORA_ROT(b, 16); \
		a ^= c; a -= ORA_ROT(c, 4); \
		b ^= a; b -= ORA_ROT(a, 14); \
		c ^= b; c -= ORA_ROT(b, 24); \
	} while (0)

/*
 * hash_bytes (src/common/hashfn.c:146-360), little-endian.  The reference has an aligned
 * word-at-a-time path and a byte-at-a-time path that produce the same value; this is the
 * byte-at-a-time form.
 */
static inline uint32_t
ora_hash_bytes(const unsigned char *k, int keylen)
{
	uint32_t	a, b, c;
	int			len = keylen;

	a = b = c = 0x9e3779b9u + (uint32_t) len + 3923095u;
	while (len >= 12)
	{
		a += (k[0] + ((uint32_t) k[1] << 8) + ((uint32_t) k[2] << 16) + ((uint32_t) k[3] << 24));
		b += (k[4] + ((uint32_t) k[5] << 8) + ((uint32_t) k[6] << 16) + ((uint32_t) k[7] << 24));
		c += (k[8] + ((uint32_t) k[9] << 8) + ((uint32_t) k[10] << 16) + ((uint32_t) k[11] << 24));
		ORA_MIX(a, b, c);
		k += 12;
		len -= 12;
	}
	switch (len)
	{
		case 11: c += ((uint32_t) k[10] << 24);	/* fall through */
		case 10: c += ((uint32_t) k[9] << 16);	/* fall through */
		case 9:  c += ((uint32_t) k[8] << 8);	/* fall through */
		/* the lowest byte of c is reserved for the length */
		case 8:  b += ((uint32_t) k[7] << 24);	/* fall through */
		case 7:  b += ((uint32_t) k[6] << 16);	/* fall through */
		case 6:  b += ((uint32_t) k[5] << 8);	/* fall through */
		case 5:  b += k[4];						/* fall through */
		case 4:  a += ((uint32_t) k[3] << 24);	/* fall through */
		case 3:  a += ((uint32_t) k[2] << 16);	/* fall through */
		case 2:  a += ((uint32_t) k[1] << 8);	/* fall through */
		case 1:  a += k[0];
	}
	ORA_FINAL(a, b, c);
	return c;
}

/* hash_bytes_uint32 (src/common/hashfn.c:627-640) = hashint4 / date hash (hashfunc.c:72) */
static inline uint32_t
ora_hash_uint32(uint32_t k)
{
	uint32_t	a, b, c;

	a = b = c = 0x9e3779b9u + (uint32_t) sizeof(uint32_t) + 3923095u;
	a += k;
	ORA_FINAL(a, b, c);
	return c;
}

/* hashint8 (src/backend/access/hash/hashfunc.c:84-102) */
static inline uint32_t
ora_hashint8(int64_t val)
{
	uint32_t	lohalf = (uint32_t) val;
	uint32_t	hihalf = (uint32_t) ((uint64_t) val >> 32);

	lohalf ^= (val >= 0) ? hihalf : ~hihalf;
	return ora_hash_uint32(lohalf);
}

/* hashfloat8 (src/backend/access/hash/hashfunc.c:194-216): +-0 -> 0, NaN canonicalised */
static inline uint32_t
ora_hashfloat8(double key)
{
	if (key == 0.0)
		return 0;
	if (key != key)
	{
		/* get_float8_nan(): the reference hashes the canonical quiet NaN bit pattern */
		uint64_t	bits = 0x7ff8000000000000ULL;
		unsigned char buf[8];

		memcpy(buf, &bits, 8);
		return ora_hash_bytes(buf, 8);
	}
	{
		unsigned char buf[8];

		memcpy(buf, &key, 8);
		return ora_hash_bytes(buf, 8);
	}
}

/* hashbpchar (src/backend/utils/adt/varchar.c:981-1004): hash_any over bytes minus trailing blanks */
static inline uint32_t
ora_hashbpchar(const char *s, int len)
{
	while (len > 0 && s[len - 1] == ' ')
		len--;
	return ora_hash_bytes((const unsigned char *) s, len);
}

/* murmurhash32 (src/include/common/hashfn.h:93-103) */
static inline uint32_t
ora_murmurhash32(uint32_t data)
{
	uint32_t	h = data;

	h ^= h >> 16;
	h *= 0x85ebca6bu;
	h ^= h >> 13;
	h *= 0xc2b2ae35u;
	h ^= h >> 16;
	return h;
}

/*
 * The per-key combine step shared by the three consumers: rotate left one bit, then XOR the
 * key's hash (NULL contributes nothing).  nodeHash.c:2134-2171 (join), execGrouping.c:473-491
 * (agg), cdbhash.c:195-217 (motion).
 */
static inline uint32_t
ora_hash_combine(uint32_t hashkey, uint32_t hkey, int isnull)
{
	hashkey = (hashkey << 1) | ((hashkey & 0x80000000u) ? 1 : 0);
	if (!isnull)
		hashkey ^= hkey;
	return hashkey;
}

/* jump_consistent_hash (src/backend/cdb/cdbhash.c:530-541) */
static inline int32_t
ora_jump_consistent_hash(uint64_t key, int32_t num_segments)
{
	int64_t		b = -1;
	int64_t		j = 0;

	while (j < num_segments)
	{
		b = j;
		key = key * 2862933555777941757ULL + 1;
		j = (int64_t) ((double) (b + 1) * ((double) (1LL << 31) / (double) ((key >> 33) + 1)));
	}
	return (int32_t) b;
}

#endif
