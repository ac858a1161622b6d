/*
 * inflate.cuh - RFC 1950 / RFC 1951 (zlib / DEFLATE) decoding for bulk-compressed AOCS blocks.
 *
 * The reference stores `compresstype=zlib` blocks (and `rle_type` compresslevel 2-4 blocks) as the output of zlib's
 * compress2() (catalog/pg_compression.c:272-318, state->compress_fn = compress2 :250) and reads them back with
 * uncompress() (:321-370) from AppendOnlyStorageRead_Content (cdb/cdbappendonlystorageread.c:1286-1310 ->
 * gp_decompress, storage/file/gp_compress.c:52-90).  zlib itself is a third-party dependency that is not under the
 * reference tree; what is restated here is the published stream format (RFC 1950 wrapper: CMF/FLG, Adler-32 trailer;
 * RFC 1951: stored / fixed / dynamic Huffman blocks, LZ77 window of 32 KiB), and the parity tests inflate streams
 * produced by the system zlib the reference would link.
 *
 * Split in two so the serial part can be tested on the host (tests/test_inflate_host.py compiles this header with g++):
 *   infl_step()  one thread: decode block headers / Huffman symbols into a queue of up to INFL_QN
 *                literal / (length, distance) entries                                  [__host__ __device__]
 *   the caller   applies the queue to the output: serially on the host, a whole warp on the device (aocs.cu)
 * (A one-stream-per-thread driver with per-lane tables interleaved in shared memory was measured and dropped: 3 warps
 * per SM could not hide the decoder's latencies, 8-30 GB/s against 22-44 GB/s for one stream per warp,
 * profiles/r01f_aocs_decode_inflate_lanes_experiment.txt.  The table stride S of the templates below is what is left.)
 */
#ifndef CB_INFLATE_CUH
#define CB_INFLATE_CUH

#include <stdint.h>

#ifdef __CUDACC__
#define INFL_HD __host__ __device__ __forceinline__
#else
#define INFL_HD static inline
#endif

#define INFL_FAST_L 10			/* first-level lookup bits, literal/length code                         */
#define INFL_FAST_D 8			/* first-level lookup bits, distance code                               */
#define INFL_QN 32				/* queue entries per step: one per lane                                 */

/* what the caller has to do after a step (always: apply the n queued entries first) */
#define INFL_MORE 0				/* call again                                                           */
#define INFL_STORED 1			/* then copy stored_len bytes from in[stored_src ...] and call again    */
#define INFL_DONE 2				/* final block finished: stream position is in `inpos` / `nbits`        */
#define INFL_ERROR 3

#define INFL_LIT 0x80000000u	/* queue entry: literal byte in bits 0-7; else length bits 0-8, distance bits 9-24 */

struct InflTables
{
	uint16_t	lfast[1 << INFL_FAST_L];	/* (symbol << 4) | code length, 0 = longer code or unused       */
	uint16_t	dfast[1 << INFL_FAST_D];
	uint16_t	lcount[16];		/* codes per length                                                     */
	uint16_t	dcount[16];
	uint16_t	lsym[288];		/* symbols in canonical order                                           */
	uint16_t	dsym[32];
	uint8_t		lens[320];		/* code lengths of the block being set up                               */
};

/* the tables as the decoder sees them: element i of a table lies at [i * S] */
template <int S>
struct InflView
{
	uint16_t   *lfast, *dfast, *lcount, *dcount, *lsym, *dsym;
	uint8_t    *lens;			/* 320 code lengths of the block being set up, stride 1                  */
	int			lbits, dbits;	/* first-level lookup bits                                               */
};

INFL_HD InflView<1>
infl_view(InflTables &T)
{
	InflView<1> V;

	V.lfast = T.lfast;
	V.dfast = T.dfast;
	V.lcount = T.lcount;
	V.dcount = T.dcount;
	V.lsym = T.lsym;
	V.dsym = T.dsym;
	V.lens = T.lens;
	V.lbits = INFL_FAST_L;
	V.dbits = INFL_FAST_D;
	return V;
}

struct InflState
{
	const uint8_t *in;
	uint32_t	inlen;
	uint32_t	inpos;			/* next byte to load into the bit buffer (may run past inlen: zeros)    */
	uint64_t	bitbuf;
	uint32_t	nbits;
	int32_t		phase;			/* 0 = at a block header, 1 = inside a Huffman block, 2 = stream finished */
	int32_t		final;
	uint32_t	stored_src;
	uint32_t	stored_len;
};

INFL_HD void
infl_init(InflState &s, const uint8_t *in, uint32_t inlen, uint32_t start)
{
	s.in = in;
	s.inlen = inlen;
	s.inpos = start;
	s.bitbuf = 0;
	s.nbits = 0;
	s.phase = 0;
	s.final = 0;
	s.stored_src = 0;
	s.stored_len = 0;
}

INFL_HD void
infl_refill(InflState &s)
{
#ifdef __CUDA_ARCH__
	/* s.in is 8-byte aligned on the device (block contents start on 8-byte file offsets): two aligned words give the
	 * next 8 stream bytes; whole bytes that fit are accounted, the rest is OR-ed in again next time (same bits) */
	if (s.inpos + 16u <= s.inlen)
	{
		const uint64_t *w = reinterpret_cast<const uint64_t *>(s.in) + (s.inpos >> 3);
		const uint32_t sh = (s.inpos & 7u) * 8u;
		uint64_t	v = w[0] >> sh;
		const uint32_t take = (63u - s.nbits) >> 3;

		if (sh)
			v |= w[1] << (64u - sh);
		s.bitbuf |= v << s.nbits;
		s.inpos += take;
		s.nbits += take * 8u;
		return;
	}
#endif
	while (s.nbits <= 56)
	{
		const uint64_t b = s.inpos < s.inlen ? s.in[s.inpos] : 0;

		s.inpos++;
		s.bitbuf |= b << s.nbits;
		s.nbits += 8;
	}
}

INFL_HD uint32_t
infl_bits(InflState &s, int n)
{
	const uint32_t v = (uint32_t) (s.bitbuf & ((1ull << n) - 1));

	s.bitbuf >>= n;
	s.nbits -= n;
	return v;
}

/* bytes of input consumed so far, counting whole bytes still in the bit buffer as not consumed */
INFL_HD uint32_t
infl_consumed(const InflState &s)
{
	return s.inpos - s.nbits / 8;
}

/* one Huffman symbol; needs 15 valid bits in the buffer.  -1 = no such code */
template <int S>
INFL_HD int
infl_sym(InflState &s, const uint16_t *fast, int fastbits, const uint16_t *count, const uint16_t *sym)
{
	const uint32_t e = fast[(s.bitbuf & ((1u << fastbits) - 1)) * S];
	uint32_t	code = 0,
				first = 0,
				index = 0;
	uint64_t	bb = s.bitbuf;

	if (e & 15)
	{
		s.bitbuf >>= (e & 15);
		s.nbits -= (e & 15);
		return (int) (e >> 4);
	}
	/* codes longer than the lookup: walk the canonical code one bit at a time */
	for (int len = 1; len <= 15; len++)
	{
		const uint32_t c = count[len * S];

		code |= (uint32_t) (bb & 1);
		bb >>= 1;
		if (code < first + c)
		{
			s.bitbuf = bb;
			s.nbits -= len;
			return sym[(index + (code - first)) * S];
		}
		index += c;
		first += c;
		first <<= 1;
		code <<= 1;
	}
	return -1;
}

/* canonical Huffman tables from code lengths; false = over-subscribed set of lengths */
template <int S>
INFL_HD bool
infl_build(const uint8_t *lens, int n, uint16_t *fast, int fastbits, uint16_t *count, uint16_t *sym)
{
	uint16_t	offs[16];
	uint16_t	next[16];
	int			left = 1;
	uint32_t	code = 0;

	for (int i = 0; i < 16; i++)
		count[i * S] = 0;
	for (int i = 0; i < n; i++)
		count[lens[i] * S]++;
	for (int len = 1; len <= 15; len++)
	{
		left <<= 1;
		left -= count[len * S];
		if (left < 0)
			return false;
	}
	offs[1] = 0;
	for (int len = 1; len < 15; len++)
		offs[len + 1] = offs[len] + count[len * S];
	for (int i = 0; i < n; i++)
		if (lens[i])
			sym[(offs[lens[i]]++) * S] = (uint16_t) i;
	for (int i = 0; i < (1 << fastbits); i++)
		fast[i * S] = 0;
	next[0] = 0;
	for (int len = 1; len <= 15; len++)
	{
		code = (code + (len > 1 ? count[(len - 1) * S] : 0)) << 1;
		next[len] = (uint16_t) code;
	}
	for (int i = 0; i < n; i++)
	{
		const int	l = lens[i];

		if (l && l <= fastbits)
		{
			uint32_t	c = next[l]++;
			uint32_t	rev = 0;

			for (int b = 0; b < l; b++)
			{
				rev = (rev << 1) | (c & 1);
				c >>= 1;
			}
			for (uint32_t j = rev; j < (1u << fastbits); j += 1u << l)
				fast[j * S] = (uint16_t) ((i << 4) | l);
		}
		else if (l)
			next[l]++;
	}
	return true;
}

/* block header at the current position: sets phase / stored_*; returns INFL_MORE / INFL_STORED / INFL_ERROR */
template <int S>
INFL_HD int
infl_block_header(InflState &s, const InflView<S> &T)
{
	uint32_t	type;

	if (infl_consumed(s) > s.inlen)
		return INFL_ERROR;
	infl_refill(s);
	s.final = (int32_t) infl_bits(s, 1);
	type = infl_bits(s, 2);
	if (type == 0)
	{
		uint32_t	len,
					nlen;

		infl_bits(s, (int) (s.nbits & 7));	/* to the byte boundary */
		infl_refill(s);
		len = infl_bits(s, 16);
		nlen = infl_bits(s, 16);
		if ((len ^ 0xFFFFu) != nlen)
			return INFL_ERROR;
		s.stored_src = infl_consumed(s);
		s.stored_len = len;
		if (s.stored_src > s.inlen || len > s.inlen - s.stored_src)
			return INFL_ERROR;
		s.inpos = s.stored_src + len;
		s.bitbuf = 0;
		s.nbits = 0;
		s.phase = s.final ? 2 : 0;
		return INFL_STORED;
	}
	if (type == 1)
	{
		for (int i = 0; i < 288; i++)
			T.lens[i] = (uint8_t) (i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8);
		infl_build<S>(T.lens, 288, T.lfast, T.lbits, T.lcount, T.lsym);
		for (int i = 0; i < 32; i++)
			T.lens[i] = 5;
		infl_build<S>(T.lens, 32, T.dfast, T.dbits, T.dcount, T.dsym);
		s.phase = 1;
		return INFL_MORE;
	}
	if (type == 2)
	{
		const uint32_t hlit = infl_bits(s, 5) + 257;
		const uint32_t hdist = infl_bits(s, 5) + 1;
		const uint32_t hclen = infl_bits(s, 4) + 4;
		uint8_t		cl[19];
		uint32_t	i = 0;

		if (hlit > 286 || hdist > 30)
			return INFL_ERROR;
		for (int k = 0; k < 19; k++)
			cl[k] = 0;
		for (uint32_t k = 0; k < hclen; k++)
		{
			/* order of the code-length code lengths: 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15 */
			const int	pos = k < 3 ? 16 + (int) k : k == 3 ? 0 : (k & 1) ? 7 - (int) ((k - 5) >> 1) : 8 + (int) ((k - 4) >> 1);

			infl_refill(s);
			cl[pos] = (uint8_t) infl_bits(s, 3);
		}
		/* the code-length code borrows the literal tables (at most 7-bit codes) */
		if (!infl_build<S>(cl, 19, T.lfast, 7, T.lcount, T.lsym))
			return INFL_ERROR;
		while (i < hlit + hdist)
		{
			int			sy;

			infl_refill(s);
			sy = infl_sym<S>(s, T.lfast, 7, T.lcount, T.lsym);
			if (sy < 0)
				return INFL_ERROR;
			if (sy < 16)
				T.lens[i++] = (uint8_t) sy;
			else
			{
				uint32_t	rep;
				uint8_t		v = 0;

				if (sy == 16)
				{
					if (i == 0)
						return INFL_ERROR;
					v = T.lens[i - 1];
					rep = 3 + infl_bits(s, 2);
				}
				else if (sy == 17)
					rep = 3 + infl_bits(s, 3);
				else
					rep = 11 + infl_bits(s, 7);
				if (i + rep > hlit + hdist)
					return INFL_ERROR;
				while (rep--)
					T.lens[i++] = v;
			}
		}
		if (T.lens[256] == 0)
			return INFL_ERROR;		/* no end-of-block code */
		if (!infl_build<S>(T.lens + hlit, (int) hdist, T.dfast, T.dbits, T.dcount, T.dsym))
			return INFL_ERROR;
		if (!infl_build<S>(T.lens, (int) hlit, T.lfast, T.lbits, T.lcount, T.lsym))
			return INFL_ERROR;
		s.phase = 1;
		return INFL_MORE;
	}
	return INFL_ERROR;
}

/* length and distance of the match whose length symbol is sy (257..285); false = bad code */
template <int S>
INFL_HD bool
infl_lendist(InflState &s, const InflView<S> &T, int sy, uint32_t *lenp, uint32_t *distp)
{
	uint32_t	len,
				dist;
	int			ds;

	sy -= 257;
	if (sy >= 29)
		return false;
	if (sy < 8)
		len = 3 + (uint32_t) sy;
	else if (sy == 28)
		len = 258;
	else
	{
		const int	ext = (sy - 4) >> 2;

		len = 3 + ((4u + ((uint32_t) sy & 3u)) << ext) + infl_bits(s, ext);
	}
	ds = infl_sym<S>(s, T.dfast, T.dbits, T.dcount, T.dsym);
	if (ds < 0 || ds >= 30)
		return false;
	if (ds < 4)
		dist = 1 + (uint32_t) ds;
	else
	{
		const int	ext = (ds - 2) >> 1;

		dist = 1 + ((2u + ((uint32_t) ds & 1u)) << ext) + infl_bits(s, ext);
	}
	*lenp = len;
	*distp = dist;
	return true;
}

/*
 * Decode until the queue holds INFL_QN entries, a block ends in a way the caller must act on, or the stream ends.
 * *n = entries queued.
 */
template <int S>
INFL_HD int
infl_step(InflState &s, const InflView<S> &T, uint32_t *q, int *n)
{
	int			cnt = 0;

	*n = 0;
	for (;;)
	{
		if (s.phase == 2)
		{
			*n = cnt;
			return INFL_DONE;
		}
		if (s.phase == 0)
		{
			const int	rc = infl_block_header<S>(s, T);

			if (rc != INFL_MORE)
			{
				*n = cnt;
				return rc;
			}
		}
		while (cnt < INFL_QN)
		{
			int			sy;

			infl_refill(s);
			sy = infl_sym<S>(s, T.lfast, T.lbits, T.lcount, T.lsym);
			if (sy < 0)
				return INFL_ERROR;
			if (sy < 256)
				q[cnt++] = INFL_LIT | (uint32_t) sy;
			else if (sy == 256)
			{
				s.phase = s.final ? 2 : 0;
				break;
			}
			else
			{
				uint32_t	len,
							dist;

				if (!infl_lendist<S>(s, T, sy, &len, &dist))
					return INFL_ERROR;
				q[cnt++] = len | (dist << 9);
			}
		}
		if (cnt == INFL_QN)
		{
			*n = cnt;
			return INFL_MORE;
		}
		if (infl_consumed(s) > s.inlen)
			return INFL_ERROR;
	}
}

/* RFC 1950 header at in[0..1]: deflate, window <= 32 KiB, no preset dictionary, check bits */
INFL_HD bool
infl_zlib_header_ok(const uint8_t *in, uint32_t inlen)
{
	if (inlen < 6)
		return false;
	return (in[0] & 0x0F) == 8 && (in[0] >> 4) <= 7 && (in[1] & 0x20) == 0 && ((uint32_t) in[0] * 256u + in[1]) % 31u == 0;
}

#endif							/* CB_INFLATE_CUH */
