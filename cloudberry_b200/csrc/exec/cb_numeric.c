/*
 * cb_numeric.c - host-side finalisation of exact aggregate states into the reference's numeric text.
 *
 * The GPU delivers, per group, an exact integer pair (sum scaled by 10^dscale as a 128-bit value,
 * N).  What the reference would print is decided by numeric_sum / numeric_avg
 * (backend/utils/adt/numeric.c:6091,6056): sum keeps the inputs' display scale; avg is
 * numeric_div(sumX, N) whose result scale comes from select_div_scale (numeric.c:9194-9254,
 * NBASE = 10000 weights and first digits) and whose last digit is rounded half away from zero
 * (div_var(..., round = true), numeric.c:2886).
 *
 * Implementation: limb arithmetic in base 10^9 (independent of the oracle's digit-string version).
 */
#include <stdio.h>
#include <string.h>

#include "../../../include/cb_exec.h"

typedef unsigned __int128 u128;
typedef __int128 i128;

#define LIMB 1000000000u
#define MAXLIMBS 24

typedef struct BigDec
{
	uint32_t	l[MAXLIMBS];	/* little endian base 1e9                                           */
	int			n;
} BigDec;

static void
big_from_u128(BigDec *b, u128 v)
{
	b->n = 0;
	memset(b->l, 0, sizeof(b->l));
	while (v)
	{
		b->l[b->n++] = (uint32_t) (v % LIMB);
		v /= LIMB;
	}
}

static void
big_mul_small(BigDec *b, uint32_t m)
{
	uint64_t	carry = 0;

	for (int i = 0; i < b->n; i++)
	{
		uint64_t	t = (uint64_t) b->l[i] * m + carry;

		b->l[i] = (uint32_t) (t % LIMB);
		carry = t / LIMB;
	}
	while (carry && b->n < MAXLIMBS)
	{
		b->l[b->n++] = (uint32_t) (carry % LIMB);
		carry /= LIMB;
	}
}

static void
big_mul_pow10(BigDec *b, int k)
{
	while (k >= 9)
	{
		big_mul_small(b, LIMB);
		k -= 9;
	}
	if (k > 0)
	{
		uint32_t	m = 1;

		while (k--)
			m *= 10;
		big_mul_small(b, m);
	}
}

/* b = floor(b / d), returns the remainder (d < 2^63) */
static uint64_t
big_div_u64(BigDec *b, uint64_t d)
{
	u128		rem = 0;

	for (int i = b->n - 1; i >= 0; i--)
	{
		u128		cur = rem * LIMB + b->l[i];

		b->l[i] = (uint32_t) (cur / d);
		rem = cur % d;
	}
	while (b->n > 0 && b->l[b->n - 1] == 0)
		b->n--;
	return (uint64_t) rem;
}

static void
big_add_one(BigDec *b)
{
	int			i = 0;

	for (;;)
	{
		if (i == b->n)
		{
			b->l[b->n++] = 1;
			return;
		}
		if (++b->l[i] < LIMB)
			return;
		b->l[i++] = 0;
	}
}

static int
big_to_digits(const BigDec *b, char *out)
{
	int			p = 0;

	if (b->n == 0)
	{
		out[p++] = '0';
		out[p] = 0;
		return p;
	}
	p += sprintf(out + p, "%u", b->l[b->n - 1]);
	for (int i = b->n - 2; i >= 0; i--)
		p += sprintf(out + p, "%09u", b->l[i]);
	return p;
}

/* digits (no point) + dscale -> numeric_out style text */
static void
put_decimal(const char *digits, int nd, int dscale, int neg, char *out, int outlen)
{
	char		buf[280];
	int			p = 0;
	int			intd = nd - dscale;

	if (neg)
		buf[p++] = '-';
	if (intd <= 0)
	{
		buf[p++] = '0';
		if (dscale > 0)
		{
			buf[p++] = '.';
			for (int i = 0; i < -intd; i++)
				buf[p++] = '0';
			memcpy(buf + p, digits, (size_t) nd);
			p += nd;
		}
	}
	else
	{
		memcpy(buf + p, digits, (size_t) intd);
		p += intd;
		if (dscale > 0)
		{
			buf[p++] = '.';
			memcpy(buf + p, digits + intd, (size_t) (nd - intd));
			p += nd - intd;
		}
	}
	buf[p] = 0;
	snprintf(out, (size_t) outlen, "%s", buf);
}

static u128
to_u128_abs(int64_t lo, int64_t hi, int *neg)
{
	i128		v = (i128) (((u128) (uint64_t) hi << 64) | (uint64_t) lo);

	*neg = v < 0;
	return v < 0 ? (u128) 0 - (u128) v : (u128) v;
}

void
cb_numeric_sum_text(int64_t lo, int64_t hi, int32_t dscale, char *out, int32_t outlen)
{
	int			neg;
	u128		a = to_u128_abs(lo, hi, &neg);
	BigDec		b;
	char		digits[256];
	int			nd;

	big_from_u128(&b, a);
	nd = big_to_digits(&b, digits);
	put_decimal(digits, nd, dscale, neg && a != 0, out, outlen);
}

static int
ndigits_u128(u128 v)
{
	int			n = 0;

	while (v)
	{
		n++;
		v /= 10;
	}
	return n;
}

static u128
pow10_u128(int k)
{
	u128		r = 1;

	while (k-- > 0)
		r *= 10;
	return r;
}

/* weight (NBASE = 10000) and first NBASE digit of a * 10^-ds, a > 0: what select_div_scale reads
 * from a normalised NumericVar */
static void
nbase_head(u128 a, int ds, int *weight, int *first)
{
	int			nd = ndigits_u128(a);
	int			intd = nd - ds;

	if (a == 0)
	{
		*weight = 0;
		*first = 0;
		return;
	}
	if (intd > 0)
	{
		int			w = (intd - 1) / 4;
		int			lead = intd - 4 * w;	/* 1..4 decimal digits in the leading group */

		*weight = w;
		*first = (int) (a / pow10_u128(nd - lead));
	}
	else
	{
		int			z = -intd;		/* zeros between the point and the first digit */
		int			take = 4 - z % 4;	/* digits of a that fall into the first non-zero group */

		*weight = -1 - z / 4;
		if (nd >= take)
			*first = (int) (a / pow10_u128(nd - take));
		else
			*first = (int) (a * pow10_u128(take - nd));
	}
}

void
cb_numeric_avg_text(int64_t lo, int64_t hi, int32_t dscale, int64_t n, char *out, int32_t outlen)
{
	int			neg,
				w1, f1, w2, f2,
				qweight,
				rscale;
	u128		a = to_u128_abs(lo, hi, &neg);
	uint64_t	dn = n < 0 ? (uint64_t) 0 - (uint64_t) n : (uint64_t) n;
	BigDec		q;
	uint64_t	rem;
	char		digits[256];
	int			nd;

	if (n == 0)
	{
		snprintf(out, (size_t) outlen, "NULL");
		return;
	}
	if (n < 0)
		neg = !neg;
	nbase_head(a, dscale, &w1, &f1);
	nbase_head((u128) dn, 0, &w2, &f2);
	qweight = w1 - w2;
	if (f1 <= f2)
		qweight--;
	rscale = 16 - qweight * 4;	/* NUMERIC_MIN_SIG_DIGITS - qweight * DEC_DIGITS */
	if (rscale < dscale)
		rscale = dscale;
	if (rscale < 0)
		rscale = 0;
	if (rscale > 1000)
		rscale = 1000;
	if (rscale - dscale > 150)
	{
		snprintf(out, (size_t) outlen, "ERR:avg scale");
		return;
	}
	big_from_u128(&q, a);
	big_mul_pow10(&q, rscale - dscale);
	rem = big_div_u64(&q, dn);
	if ((u128) rem * 2 >= (u128) dn)
		big_add_one(&q);
	nd = big_to_digits(&q, digits);
	put_decimal(digits, nd, rscale, neg && q.n != 0, out, outlen);
}
