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

/* ------------------------------------------------------------------------------------------
 * Partial aggregate states in the reference's serialised form.
 *
 * When a Partial Aggregate runs on the device and its Finalize stage on a CPU process (or the other way round), the
 * state crosses the Motion as the `bytea` the aggregate's serialisation function makes (AGGSPLIT_INITIAL_SERIAL /
 * FINAL_DESERIAL, nodes/nodes.h:977-1000):
 *
 *   numeric_avg_serialize (utils/adt/numeric.c:5025-5080; sum / avg over numeric):
 *       int64 N | numeric_send(sumX) | int32 maxScale | int64 maxScaleCount | int64 NaNcount | int64 pInfcount | int64 nInfcount
 *   int8_avg_serialize (numeric.c:5793; sum / avg over int8 with 128-bit states):
 *       int64 N | numeric_send(sumX)
 *   numeric_send (numeric.c:1069): int16 ndigits | int16 weight | int16 sign | int16 dscale | ndigits x int16 base-10000 digits
 *
 * all big-endian (pq_send*).  The device state (N, exact 128-bit sum scaled by 10^dscale) determines every field: the
 * columns on this path have ONE display scale, so maxScale = dscale and maxScaleCount = N (numeric.c:4690-4705
 * do_numeric_accum keeps the count of inputs at the largest scale seen), and NaN / infinity counts are 0 - the device has no
 * such values.  sumX is normalised as make_result does (strip_var: no leading / trailing zero digits, weight 0 for zero).
 * ------------------------------------------------------------------------------------------ */
static int
put_be16(uint8_t *out, int at, int v)
{
	out[at] = (uint8_t) ((unsigned) v >> 8);
	out[at + 1] = (uint8_t) v;
	return at + 2;
}

static int
put_be64(uint8_t *out, int at, int64_t v)
{
	for (int i = 0; i < 8; i++)
		out[at + i] = (uint8_t) ((uint64_t) v >> (56 - 8 * i));
	return at + 8;
}

/* numeric_send of (hi:lo) * 10^-dscale; returns the new offset, or -1 when it does not fit `cap` */
static int
numeric_send_i128(int64_t lo, int64_t hi, int32_t dscale, uint8_t *out, int at, int cap)
{
	i128		sv = (i128) ((((u128) (uint64_t) hi) << 64) | (u128) (uint64_t) lo);
	const int	neg = sv < 0;
	u128		v = neg ? (u128) (-sv) : (u128) sv;
	const int	groups_after = (dscale + 3) / 4;
	uint16_t	rev[16],
				digits[16];
	int			k = 0,
				nd,
				first = 0,
				weight;
	BigDec		b;

	/* align the decimal point to a base-10000 digit boundary: value * 10^(4 * groups_after - dscale), in limbs (the product
	 * can leave 128 bits) */
	big_from_u128(&b, v);
	for (int i = 0; i < 4 * groups_after - dscale; i++)
		big_mul_small(&b, 10);
	{
		/* base 1e9 limbs -> decimal string -> groups of four from the right */
		char		dec[MAXLIMBS * 9 + 2];
		int			len = 0;

		if (b.n == 0)
			dec[len++] = '0';
		else
		{
			len += snprintf(dec + len, sizeof(dec) - (size_t) len, "%u", b.l[b.n - 1]);
			for (int i = b.n - 2; i >= 0; i--)
				len += snprintf(dec + len, sizeof(dec) - (size_t) len, "%09u", b.l[i]);
		}
		for (int end = len; end > 0 && k < 16; end -= 4)
		{
			int			start = end - 4 < 0 ? 0 : end - 4;
			int			g = 0;

			for (int i = start; i < end; i++)
				g = g * 10 + (dec[i] - '0');
			rev[k++] = (uint16_t) g;
		}
	}
	nd = k;
	for (int i = 0; i < k; i++)
		digits[i] = rev[k - 1 - i];
	weight = nd - groups_after - 1;
	while (first < nd && digits[first] == 0)
	{
		first++;
		weight--;
	}
	while (nd > first && digits[nd - 1] == 0)
		nd--;
	if (nd == first)
		weight = 0;				/* strip_var: zero has weight 0 and no digits */
	if (at + 8 + 2 * (nd - first) > cap)
		return -1;
	at = put_be16(out, at, nd - first);
	at = put_be16(out, at, weight);
	at = put_be16(out, at, (nd > first && neg) ? 0x4000 : 0x0000);	/* NUMERIC_NEG / NUMERIC_POS */
	at = put_be16(out, at, dscale);
	for (int i = first; i < nd; i++)
		at = put_be16(out, at, digits[i]);
	return at;
}

int
cb_numeric_avg_serialize(int64_t n, int64_t sum_lo, int64_t sum_hi, int32_t dscale, uint8_t *out, int32_t cap)
{
	int			at = 0;

	if (cap < 8 || dscale < 0 || dscale > 0x3FFF || n < 0)
		return -1;
	at = put_be64(out, at, n);
	at = numeric_send_i128(sum_lo, sum_hi, dscale, out, at, cap);
	if (at < 0 || at + 4 + 4 * 8 > cap)
		return -1;
	/* maxScale: do_numeric_accum records the largest display scale seen, 0 before the first input */
	out[at++] = 0;
	out[at++] = 0;
	at = put_be16(out, at, n > 0 ? dscale : 0);
	at = put_be64(out, at, n > 0 ? n : 0);	/* maxScaleCount */
	at = put_be64(out, at, 0);				/* NaNcount */
	at = put_be64(out, at, 0);				/* pInfcount */
	at = put_be64(out, at, 0);				/* nInfcount */
	return at;
}

int
cb_int8_avg_serialize(int64_t n, int64_t sum_lo, int64_t sum_hi, uint8_t *out, int32_t cap)
{
	int			at = 0;

	if (cap < 8 || n < 0)
		return -1;
	at = put_be64(out, at, n);
	at = numeric_send_i128(sum_lo, sum_hi, 0, out, at, cap);
	return at;
}

static int64_t
get_be64(const uint8_t *in)
{
	uint64_t	v = 0;

	for (int i = 0; i < 8; i++)
		v = (v << 8) | in[i];
	return (int64_t) v;
}

static int
get_be16(const uint8_t *in)
{
	return (int) (int16_t) (((unsigned) in[0] << 8) | in[1]);
}

/* the reverse (numeric_avg_deserialize numeric.c:5092 / int8_avg_deserialize): a state serialised by a CPU Partial Aggregate
 * -> (N, 128-bit sum at display scale *dscale).  with_tail: numeric_avg_serialize's trailing fields are present and checked.
 * Returns 0, or < 0: -1 malformed, -2 NaN / infinity inputs (no device representation), -3 the sum leaves 128 bits or carries
 * more fractional digits than its display scale. */
int
cb_numeric_avg_deserialize(const uint8_t *in, int32_t len, int32_t with_tail, int64_t *n, int64_t *sum_lo, int64_t *sum_hi, int32_t *dscale)
{
	int			nd,
				weight,
				sign,
				ds;
	i128		v = 0;

	if (len < 16)
		return -1;
	*n = get_be64(in);
	nd = get_be16(in + 8);
	weight = get_be16(in + 10);
	sign = get_be16(in + 12) & 0xFFFF;
	ds = get_be16(in + 14);
	if (nd < 0 || 16 + 2 * nd + (with_tail ? 36 : 0) != len || ds < 0)
		return -1;
	if (sign != 0x0000 && sign != 0x4000)
		return -2;				/* NUMERIC_NAN / infinities */
	if (with_tail)
	{
		const uint8_t *t = in + 16 + 2 * nd;

		if (get_be64(t + 12) != 0 || get_be64(t + 20) != 0 || get_be64(t + 28) != 0)
			return -2;
	}
	/* value = sum digit[i] * 10000^(weight - i); scaled = value * 10^ds must be an integer that fits 127 bits */
	for (int i = 0; i < nd; i++)
	{
		const int	e = 4 * (weight - i) + ds;
		i128		d = get_be16(in + 16 + 2 * i);

		if (d < 0 || d > 9999)
			return -1;
		if (e < 0)
		{
			i128		div = 1;

			if (-e > 4)
			{
				if (d)
					return -3;
				continue;
			}
			for (int k = 0; k < -e; k++)
				div *= 10;
			if (d % div)
				return -3;
			d /= div;
		}
		else
		{
			if (e > 38)
				return -3;
			for (int k = 0; k < e; k++)
			{
				if (d > ((((i128) 1) << 126) / 10))
					return -3;
				d *= 10;
			}
		}
		if (v > ((((i128) 1) << 126)) - d)
			return -3;
		v += d;
	}
	if (sign == 0x4000)
		v = -v;
	*sum_lo = (int64_t) (uint64_t) (u128) v;
	*sum_hi = (int64_t) (uint64_t) (((u128) v) >> 64);
	*dscale = ds;
	return 0;
}
