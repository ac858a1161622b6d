/*
 * cb_tupser.c - rows of the GPU executor in the reference's Motion wire format, and back.
 *
 * a MotionIPCLayer-compatible interconnect lets un-replaced CPU operators exchange tuples with a GPU
 * segment.  What travels in the reference's interconnect packets is a stream of tuple chunks
 * (cdb/motion/tupser.c:349-513 SerializeTuple, :515-690 CvtChunksToTup; include/cdb/tupchunk.h): a 4-byte chunk header
 * (uint16 payload size, uint16 TupleChunkType, host byte order), and as payload an int32 length followed by the body of
 * a MinimalTuple from t_infomask2 on (MINIMAL_TUPLE_DATA_OFFSET, include/access/htup_details.h), cut into chunks of at
 * most Gp_max_tuple_chunk_size bytes (TC_WHOLE, or TC_PARTIAL_START / _MID / _END); TC_END_OF_STREAM closes a sender.
 * The MinimalTuple body is what heap_form_minimal_tuple lays out (access/common/heaptuple.c:1453-1510, heap_fill_tuple
 * :304-400): natts and flags, t_hoff, the NULL bitmap when some attribute is NULL, then each attribute at its type's
 * alignment counted from the start of the (virtual) heap tuple header -- except varlenas that fit a 1-byte header,
 * which are packed unaligned (fill_val :220-290, VARATT_CAN_MAKE_SHORT).
 *
 * Column kinds here are the executor's (cb_plan.h CbTypeId): integers, date, float8, bool by value; numeric from the
 * scaled int64 the device computes with back to the reference's base-10000 digit string (make_result / strip_var,
 * utils/adt/numeric.c); character(1) as the 1-byte string it is; dictionary columns as the text the code stands for
 * (blank-padded to the declared width for character(n), as bpchar values always are).
 *
 * Host code on purpose: a Motion towards a CPU operator ends in host packets anyway; the device part of such a Motion
 * is the partitioning (k_pipeline / k_probe_chain PARTITION sinks), the rows arrive here through cbgpu_rel_read_rows.
 */
#include "../../../include/cb_exec.h"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TC_HDR 4				/* TUPLE_CHUNK_HEADER_SIZE */
enum
{
	TC_WHOLE, TC_PARTIAL_START, TC_PARTIAL_MID, TC_PARTIAL_END, TC_END_OF_STREAM, TC_EMPTY
};

#define HEAP_HASNULL 0x0001
#define HEAP_HASVARWIDTH 0x0002
/* offsetof(HeapTupleHeaderData, t_bits) = 23; a MinimalTuple body starts at header offset 18 (t_infomask2) */
#define HTUP_BITS_OFF 23
#define HTUP_BODY_OFF 18
#define MAXALIGN8(x) (((x) + 7) & ~7)

static int
attr_is_varlena(int32_t type)
{
	return type == CB_NUMERIC || type == CB_BPCHAR1 || type == CB_DICT8 || type == CB_DICT32 || type == CB_TUPSER_STATE_NUMERIC ||
		type == CB_TUPSER_STATE_INT8;
}

/* bytes of the attribute's by-value datum, and its alignment */
static int
attr_fixed_len(int32_t type)
{
	switch (type)
	{
		case CB_INT4: case CB_DATE: return 4;
		case CB_INT8: case CB_FLOAT8: return 8;
		case CB_BOOL: return 1;
		default: return -1;
	}
}

/*
 * scaled * 10^-dscale as a numeric datum body (after the varlena header): NumericShort when it fits, else NumericLong
 * (utils/adt/numeric.c make_result :7520-7590, NUMERIC_CAN_BE_SHORT).  Returns the length.
 */
static int
numeric_body(int64_t scaled, int32_t dscale, unsigned char *out)
{
	unsigned short digits[8];
	int			nd = 0,
				weight,
				first = 0,
				n = 0;
	const int	groups_after = (dscale + 3) / 4;
	const int	neg = scaled < 0;
	unsigned __int128 v = neg ? (unsigned __int128) (-(__int128) scaled) : (unsigned __int128) scaled;

	for (int i = 0; i < 4 * groups_after - dscale; i++)
		v *= 10;
	{
		unsigned short rev[12];
		int			k = 0;

		while (v)
		{
			rev[k++] = (unsigned short) (v % 10000);
			v /= 10000;
		}
		nd = k;
		for (int i = 0; i < k; i++)
			digits[i] = rev[k - 1 - i];
	}
	weight = nd - groups_after - 1;
	while (first < nd && digits[first] == 0)
	{
		first++;
		weight--;
	}
	while (nd > first && digits[nd - 1] == 0)
		nd--;
	if (nd == first)
		weight = 0;
	{
		const int	sign = (nd > first) && neg;

		if (dscale <= 0x3F && weight >= -64 && weight <= 63)
		{
			const unsigned hdr = 0x8000u | (sign ? 0x2000u : 0u) | ((unsigned) dscale << 7) | (weight < 0 ? 0x0040u : 0u) | ((unsigned) weight & 0x003Fu);

			out[n++] = (unsigned char) hdr;
			out[n++] = (unsigned char) (hdr >> 8);
		}
		else
		{
			const unsigned hdr = (sign ? 0x4000u : 0u) | ((unsigned) dscale & 0x3FFFu);

			out[n++] = (unsigned char) hdr;
			out[n++] = (unsigned char) (hdr >> 8);
			out[n++] = (unsigned char) weight;
			out[n++] = (unsigned char) ((unsigned) weight >> 8);
		}
	}
	for (int i = first; i < nd; i++)
	{
		out[n++] = (unsigned char) digits[i];
		out[n++] = (unsigned char) (digits[i] >> 8);
	}
	return n;
}

/* the bytes of a varlena attribute's value (no header) */
static int
varlena_payload(const CbTupAttr *a, int64_t v, unsigned char *scratch, const unsigned char **data, int *len)
{
	if (a->type == CB_NUMERIC)
	{
		*len = numeric_body(v, a->dscale, scratch);
		*data = scratch;
		return 0;
	}
	if (a->type == CB_BPCHAR1)
	{
		scratch[0] = (unsigned char) v;
		*data = scratch;
		*len = 1;
		return 0;
	}
	if (a->type == CB_TUPSER_STATE_NUMERIC || a->type == CB_TUPSER_STATE_INT8)
	{
		/* a partial aggregate state: the bytea its serialisation function would make (numeric.c:5025 / :5793) */
		const CbAggStateDatum *st = (const CbAggStateDatum *) (intptr_t) v;
		const int	n = st == NULL ? -1 : a->type == CB_TUPSER_STATE_NUMERIC
			? cb_numeric_avg_serialize(st->n, st->lo, st->hi, a->dscale, scratch, 256)
			: cb_int8_avg_serialize(st->n, st->lo, st->hi, scratch, 256);

		if (n < 0)
			return -1;
		*data = scratch;
		*len = n;
		return 0;
	}
	{
		const char *text;
		int32_t		tl;

		if (v < 0 || v >= a->ntexts)
			return -1;
		text = a->texts[v];
		tl = a->text_lens[v];
		if (a->bpchar_len > tl)
		{
			/* character(n): stored blank-padded to n (bpchar_input, utils/adt/varchar.c:130-200) */
			if (a->bpchar_len > CB_TUPSER_MAX_TEXT)
				return -1;
			memcpy(scratch, text, (size_t) tl);
			memset(scratch + tl, ' ', (size_t) (a->bpchar_len - tl));
			*data = scratch;
			*len = a->bpchar_len;
		}
		else
		{
			*data = (const unsigned char *) text;
			*len = tl;
		}
	}
	return 0;
}

int64_t
cb_tupser_row(const CbTupAttr *attrs, int natts, const int64_t *values, const uint8_t *isnull, int max_chunk, unsigned char *out,
			  int64_t outcap)
{
	unsigned char scratch[CB_TUPSER_MAX_TEXT + 64];
	int			hasnull = 0,
				hasvar = 0;
	int			hoff,
				off;
	int64_t		need;
	unsigned char *body;
	int			bodylen;

	if (natts < 0 || natts > CB_TUPSER_MAX_ATTS || max_chunk <= TC_HDR + 4)
		return -1;
	/* (a row without attributes goes the same way: a 6-byte body; the TC_EMPTY shortcut of the direct transport buffer,
	 * tupser.c:366-373, is accepted by cb_tupser_next but not produced) */
	for (int i = 0; i < natts; i++)
	{
		if (isnull && isnull[i])
			hasnull = 1;
		if (attr_is_varlena(attrs[i].type))
			hasvar |= !(isnull && isnull[i]);	/* HEAP_HASVARWIDTH is set by fill_val when it stores one (heaptuple.c:240) */
		else if (attr_fixed_len(attrs[i].type) < 0)
			return -1;
	}
	hoff = MAXALIGN8(HTUP_BITS_OFF + (hasnull ? (natts + 7) / 8 : 0));
	/* pass 1: length (heap_compute_data_size, heaptuple.c:120-200) */
	off = hoff;
	for (int i = 0; i < natts; i++)
	{
		if (isnull && isnull[i])
			continue;
		if (attr_is_varlena(attrs[i].type))
		{
			const unsigned char *d;
			int			l;

			if (varlena_payload(&attrs[i], values[i], scratch, &d, &l) != 0)
				return -1;
			if (l + 1 <= 0x7F)
				off += l + 1;	/* 1-byte header, no alignment */
			else
				off = ((off + 3) & ~3) + 4 + l;
		}
		else
		{
			const int	w = attr_fixed_len(attrs[i].type);

			off = ((off + w - 1) & ~(w - 1)) + w;
		}
	}
	bodylen = off - HTUP_BODY_OFF;
	body = calloc(1, (size_t) bodylen + 8);
	if (!body)
		return -3;
	/* pass 2: the body (heap_fill_tuple) */
	body[0] = (unsigned char) natts;				/* t_infomask2: number of attributes */
	body[1] = (unsigned char) (natts >> 8);
	body[2] = (unsigned char) ((hasnull ? HEAP_HASNULL : 0) | (hasvar ? HEAP_HASVARWIDTH : 0));	/* t_infomask */
	body[3] = 0;
	body[4] = (unsigned char) hoff;
	if (hasnull)
		for (int i = 0; i < natts; i++)
			if (!isnull[i])
				body[HTUP_BITS_OFF - HTUP_BODY_OFF + (i >> 3)] |= (unsigned char) (1 << (i & 7));	/* 1 = not null */
	off = hoff;
	for (int i = 0; i < natts; i++)
	{
		if (isnull && isnull[i])
			continue;
		if (attr_is_varlena(attrs[i].type))
		{
			const unsigned char *d;
			int			l;

			varlena_payload(&attrs[i], values[i], scratch, &d, &l);
			if (l + 1 <= 0x7F)
			{
				body[off - HTUP_BODY_OFF] = (unsigned char) (((l + 1) << 1) | 1);	/* SET_VARSIZE_SHORT */
				memcpy(body + off - HTUP_BODY_OFF + 1, d, (size_t) l);
				off += l + 1;
			}
			else
			{
				const unsigned total = (unsigned) (l + 4) << 2;	/* SET_VARSIZE, little endian */

				off = (off + 3) & ~3;
				memcpy(body + off - HTUP_BODY_OFF, &total, 4);
				memcpy(body + off - HTUP_BODY_OFF + 4, d, (size_t) l);
				off += 4 + l;
			}
		}
		else
		{
			const int	w = attr_fixed_len(attrs[i].type);

			off = (off + w - 1) & ~(w - 1);
			memcpy(body + off - HTUP_BODY_OFF, &values[i], (size_t) w);	/* little endian: the low bytes */
			off += w;
		}
	}
	/* chunks: [int32 bodylen][body] cut at max_chunk - header bytes per chunk (SerializeTuple / addByteStringToChunkList) */
	{
		const int64_t total = 4 + (int64_t) bodylen;
		const int	per = max_chunk - TC_HDR;
		const int64_t nchunks = (total + per - 1) / per;
		int64_t		done = 0,
					pos = 0;
		unsigned char lenbuf[4];

		need = total + nchunks * TC_HDR;
		if (need > outcap)
		{
			free(body);
			return -2;
		}
		memcpy(lenbuf, &bodylen, 4);
		for (int64_t c = 0; c < nchunks; c++)
		{
			const int	n = (int) (total - done < per ? total - done : per);
			const int	type = nchunks == 1 ? TC_WHOLE : c == 0 ? TC_PARTIAL_START : c == nchunks - 1 ? TC_PARTIAL_END : TC_PARTIAL_MID;

			out[pos] = (unsigned char) n;
			out[pos + 1] = (unsigned char) (n >> 8);
			out[pos + 2] = (unsigned char) type;
			out[pos + 3] = 0;
			for (int k = 0; k < n; k++)
			{
				const int64_t src = done + k;

				out[pos + TC_HDR + k] = src < 4 ? lenbuf[src] : body[src - 4];
			}
			pos += TC_HDR + n;
			done += n;
		}
	}
	free(body);
	return need;
}

int
cb_tupser_end_of_stream(unsigned char *out, int64_t outcap)
{
	if (outcap < TC_HDR)
		return -2;
	out[0] = out[1] = 0;
	out[2] = TC_END_OF_STREAM;
	out[3] = 0;
	return TC_HDR;
}

/* code of a string in the attribute's dictionary (texts in byte order, as cbgpu_dict_entry lists them), -1 when absent;
 * character(n) values compare without their trailing blanks */
static int32_t
text_code(const CbTupAttr *a, const unsigned char *d, int l)
{
	int32_t		lo = 0,
				hi = a->ntexts - 1;

	if (a->bpchar_len)
		while (l > 0 && d[l - 1] == ' ')
			l--;
	while (lo <= hi)
	{
		const int32_t mid = (lo + hi) / 2;
		const int	ml = a->text_lens[mid];
		int			c = memcmp(a->texts[mid], d, (size_t) (ml < l ? ml : l));

		if (c == 0)
			c = ml < l ? -1 : ml > l;
		if (c == 0)
			return mid;
		if (c < 0)
			lo = mid + 1;
		else
			hi = mid - 1;
	}
	return -1;
}

/* numeric datum body -> value scaled by 10^dscale; 0 when it does not fit an int64 or carries more scale */
static int
numeric_to_scaled(const unsigned char *b, int len, int32_t dscale, int64_t *out)
{
	unsigned	hdr;
	int			weight,
				nd,
				sign,
				p = 2;
	__int128	v = 0;

	if (len < 2)
		return 0;
	hdr = (unsigned) b[0] | ((unsigned) b[1] << 8);
	if ((hdr & 0xC000u) == 0xC000u)
		return 0;				/* NaN / infinities */
	if (hdr & 0x8000u)
	{
		sign = (hdr & 0x2000u) != 0;
		weight = (int) (hdr & 0x003Fu);
		if (hdr & 0x0040u)
			weight |= ~0x3F;
	}
	else
	{
		if (len < 4)
			return 0;
		sign = (hdr & 0x4000u) != 0;
		weight = (short) ((unsigned) b[2] | ((unsigned) b[3] << 8));
		p = 4;
	}
	nd = (len - p) / 2;
	/* value = sum digit[i] * 10000^(weight - i); scaled = value * 10^dscale */
	for (int i = 0; i < nd; i++)
	{
		const int	e = 4 * (weight - i) + dscale;	/* power of ten of this digit's unit */
		__int128	d = (int) ((unsigned) b[p + 2 * i] | ((unsigned) b[p + 2 * i + 1] << 8));

		if (e < 0)
		{
			/* digits below the target scale must be zero once shifted */
			int			k = -e;
			__int128	div = 1;

			if (k > 4)
			{
				if (d)
					return 0;
				continue;
			}
			while (k--)
				div *= 10;
			if (d % div)
				return 0;
			d /= div;
		}
		else
		{
			if (e > 36)
				return 0;
			for (int k = 0; k < e; k++)
				d *= 10;
		}
		v += d;
		if (v > (__int128) INT64_MAX)
			return 0;
	}
	*out = (int64_t) (sign ? -v : v);
	return 1;
}

int64_t
cb_tupser_next(const CbTupAttr *attrs, int natts, const unsigned char *in, int64_t inlen, int64_t *consumed, int64_t *values,
			   uint8_t *isnull)
{
	unsigned char *buf = NULL;
	int64_t		pos = 0,
				got = 0,
				cap = 0;
	int			state = 0;		/* 0 nothing yet, 1 inside a partial tuple */

	*consumed = 0;
	for (;;)
	{
		int			size,
					type;

		if (pos + TC_HDR > inlen)
		{
			free(buf);
			return CB_TUPSER_NEED_MORE;
		}
		size = in[pos] | (in[pos + 1] << 8);
		type = in[pos + 2] | (in[pos + 3] << 8);
		if (pos + TC_HDR + size > inlen)
		{
			free(buf);
			return CB_TUPSER_NEED_MORE;
		}
		if (type == TC_END_OF_STREAM && state == 0)
		{
			*consumed = pos + TC_HDR;
			return CB_TUPSER_END;
		}
		if (type == TC_EMPTY && state == 0)
		{
			*consumed = pos + TC_HDR;
			return natts == 0 ? 1 : CB_TUPSER_BAD;
		}
		if ((state == 0 && type != TC_WHOLE && type != TC_PARTIAL_START) || (state == 1 && type != TC_PARTIAL_MID && type != TC_PARTIAL_END))
		{
			free(buf);
			return CB_TUPSER_BAD;	/* CvtChunksToTup's protocol checks (tupser.c:538-640) */
		}
		if (got + size > cap)
		{
			cap = (got + size) * 2 + 64;
			unsigned char *grown = realloc(buf, (size_t) cap);

			if (!grown)
			{
				free(buf);
				return CB_TUPSER_BAD;
			}
			buf = grown;
		}
		if (size)
			memcpy(buf + got, in + pos + TC_HDR, (size_t) size);
		got += size;
		pos += TC_HDR + size;
		if (type == TC_WHOLE || type == TC_PARTIAL_END)
			break;
		state = 1;
	}
	*consumed = pos;
	/* [int32 bodylen][MinimalTuple body]: heap_deform_tuple (heaptuple.c:1249-1350) */
	{
		int			bodylen;
		const unsigned char *body = buf + 4;
		int			n,
					hasnull,
					hoff,
					off;

		if (got < 4 + 5)
		{
			free(buf);
			return CB_TUPSER_BAD;
		}
		memcpy(&bodylen, buf, 4);
		if (bodylen < 5 || (int64_t) bodylen + 4 != got)
		{
			free(buf);
			return CB_TUPSER_BAD;
		}
		n = (body[0] | (body[1] << 8)) & 0x07FF;	/* HEAP_NATTS_MASK */
		hasnull = body[2] & HEAP_HASNULL;
		hoff = body[4];
		if (n != natts || hoff < HTUP_BITS_OFF || hoff - HTUP_BODY_OFF > bodylen ||
			(hasnull && hoff < HTUP_BITS_OFF + (natts + 7) / 8))
		{
			/* with HEAP_HASNULL the header must cover the whole NULL bitmap: a short tuple claiming many attributes would
			 * otherwise be read beyond its body (heap_form_minimal_tuple: t_hoff = MAXALIGN(offsetof(t_bits) + BITMAPLEN)) */
			free(buf);
			return CB_TUPSER_BAD;
		}
		off = hoff;
		for (int i = 0; i < natts; i++)
		{
			values[i] = 0;
			isnull[i] = 0;
			if (hasnull && !((body[HTUP_BITS_OFF - HTUP_BODY_OFF + (i >> 3)] >> (i & 7)) & 1))
			{
				isnull[i] = 1;
				continue;
			}
			if (attr_is_varlena(attrs[i].type))
			{
				const unsigned char *d;
				int			l;

				if (off - HTUP_BODY_OFF >= bodylen)
					goto bad;
				if (body[off - HTUP_BODY_OFF] == 0)
					off = (off + 3) & ~3;	/* a pad byte: the datum has a 4-byte header (att_align_pointer) */
				if (off - HTUP_BODY_OFF >= bodylen)
					goto bad;
				if (body[off - HTUP_BODY_OFF] & 1)
				{
					l = (body[off - HTUP_BODY_OFF] >> 1) - 1;
					d = body + off - HTUP_BODY_OFF + 1;
					off += l + 1;
				}
				else
				{
					unsigned	total;

					memcpy(&total, body + off - HTUP_BODY_OFF, 4);
					if (total & 3)
						goto bad;	/* compressed / external datums do not travel (SerializeTuple detoasts) */
					l = (int) (total >> 2) - 4;
					d = body + off - HTUP_BODY_OFF + 4;
					off += 4 + l;
				}
				if (l < 0 || off - HTUP_BODY_OFF > bodylen)
					goto bad;
				if (attrs[i].type == CB_NUMERIC)
				{
					if (!numeric_to_scaled(d, l, attrs[i].dscale, &values[i]))
						goto bad;
				}
				else if (attrs[i].type == CB_BPCHAR1)
					values[i] = l > 0 ? d[0] : ' ';
				else if (attrs[i].type == CB_TUPSER_STATE_NUMERIC || attrs[i].type == CB_TUPSER_STATE_INT8)
				{
					/* a CPU Partial Aggregate's serialised state -> (N, sum): numeric_avg_deserialize / int8_avg_deserialize */
					CbAggStateDatum *st = attrs[i].state;
					int32_t		ds = 0;

					if (st == NULL || cb_numeric_avg_deserialize(d, l, attrs[i].type == CB_TUPSER_STATE_NUMERIC, &st->n, &st->lo, &st->hi, &ds) != 0)
						goto bad;
					if (ds != (attrs[i].type == CB_TUPSER_STATE_NUMERIC ? attrs[i].dscale : 0))
					{
						/* the sender's sum carries another display scale (inputs of mixed scale): bring it to ours if exact */
						if (ds > attrs[i].dscale || attrs[i].type != CB_TUPSER_STATE_NUMERIC)
							goto bad;
						{
							__int128	v = (__int128) ((((unsigned __int128) (uint64_t) st->hi) << 64) | (uint64_t) st->lo);

							for (int k = ds; k < attrs[i].dscale; k++)
							{
								if (v > (((__int128) 1) << 122) || v < -(((__int128) 1) << 122))
									goto bad;
								v *= 10;
							}
							st->lo = (int64_t) (uint64_t) (unsigned __int128) v;
							st->hi = (int64_t) (uint64_t) (((unsigned __int128) v) >> 64);
						}
					}
					values[i] = (int64_t) (intptr_t) st;
				}
				else
				{
					const int32_t code = text_code(&attrs[i], d, l);

					if (code < 0)
						goto bad;	/* a string this segment's dictionary does not hold */
					values[i] = code;
				}
			}
			else
			{
				const int	w = attr_fixed_len(attrs[i].type);

				off = (off + w - 1) & ~(w - 1);
				if (off + w - HTUP_BODY_OFF > bodylen)
					goto bad;
				if (w == 8)
					memcpy(&values[i], body + off - HTUP_BODY_OFF, 8);
				else if (w == 4)
				{
					int32_t		x;

					memcpy(&x, body + off - HTUP_BODY_OFF, 4);
					values[i] = x;
				}
				else
					values[i] = body[off - HTUP_BODY_OFF];
				off += w;
			}
		}
	}
	free(buf);
	return 1;
bad:
	free(buf);
	return CB_TUPSER_BAD;
}
