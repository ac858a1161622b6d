/*
 * zstd_dec.cuh - Zstandard frame decoding (RFC 8878) for bulk-compressed AOCS blocks.
 *
 * The reference stores `compresstype=zstd` blocks as the output of ZSTD_compressCCtx() and reads them back with
 * ZSTD_decompressDCtx() (gpcontrib/zstd/zstd_compression.c:104-140, 142-175), called from
 * AppendOnlyStorageRead_Content (cdb/cdbappendonlystorageread.c:1286-1310 -> gp_decompress).  libzstd is a
 * third-party dependency that is not vendored under the reference tree; what is restated here is the published frame
 * format: frame header, Raw / RLE / Compressed blocks, literals (Raw / RLE / Huffman in 1 or 4 streams / treeless),
 * Huffman tree descriptions (direct or FSE-compressed weights), sequences with Predefined / RLE / FSE / Repeat tables,
 * repeat-offset history, optional XXH64 content checksum.  Dictionaries are not used by the reference and are refused.
 * Parity is pinned against streams produced by a real libzstd (the one bundled with pyarrow, tests/test_zstd_host.py
 * and tests/golden/aocs_zstd_columns.npz).
 *
 * As with inflate.cuh the serial part is __host__ __device__ so it can be tested on the host; aocs.cu adds the
 * warp-parallel parts (4 Huffman streams on 4 lanes, batch execution of sequences).
 */
#ifndef CB_ZSTD_DEC_CUH
#define CB_ZSTD_DEC_CUH

#include <stdint.h>

#ifdef __CUDACC__
#define Z_HD __host__ __device__ __forceinline__
#define Z_DEVCONST __device__ static const
#else
#define Z_HD static inline
#endif

/* tables needed on both sides: one initialiser, a host and (under nvcc) a device copy */
#ifdef __CUDACC__
#define Z_TABLE(type, name, n, ...) static const type name##_h[n] = __VA_ARGS__; Z_DEVCONST type name##_d[n] = __VA_ARGS__;
#else
#define Z_TABLE(type, name, n, ...) static const type name##_h[n] = __VA_ARGS__;
#endif
#ifdef __CUDA_ARCH__
#define Z_T(name) name##_d
#else
#define Z_T(name) name##_h
#endif

#define Z_BLOCK_MAX (128u * 1024u)	/* Block_Maximum_Size                                                   */
#define Z_LL_MAXLOG 9
#define Z_OF_MAXLOG 8
#define Z_ML_MAXLOG 9
#define Z_HUF_MAXLOG 11
#define Z_SEQ_QN 32

/* RFC 8878 3.1.1.3.2.1.1: sequence code -> baseline, extra bits */
Z_TABLE(uint32_t, z_ll_base, 36, {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256,
								  512, 1024, 2048, 4096, 8192, 16384, 32768, 65536})
Z_TABLE(uint8_t, z_ll_bits, 36, {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16})
Z_TABLE(uint32_t, z_ml_base, 53, {3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31,
								  32, 33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771,
								  65539})
Z_TABLE(uint8_t, z_ml_bits, 53, {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1,
								 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16})
/* 3.1.1.3.2.2: default distributions (accuracy 6, 6, 5) */
Z_TABLE(int16_t, z_ll_default, 36, {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1,
									-1, -1})
Z_TABLE(int16_t, z_ml_default, 53, {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
									1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1})
Z_TABLE(int16_t, z_of_default, 29, {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1})

struct ZFse
{
	uint8_t		sym;
	uint8_t		nbits;
	uint16_t	base;
};

/* decoding tables of one frame: kept between the blocks of the frame (Repeat modes, treeless literals) */
struct ZTab
{
	ZFse		ll[1 << Z_LL_MAXLOG];
	ZFse		of[1 << Z_OF_MAXLOG];
	ZFse		ml[1 << Z_ML_MAXLOG];
	ZFse		wfse[64];		/* FSE table of a compressed Huffman weight list                       */
	uint16_t	huf[1 << Z_HUF_MAXLOG];	/* symbol | number of bits << 8                                */
	int16_t		norm[64];		/* scratch: normalised counts                                           */
	uint16_t	next[64];		/* scratch: next state per symbol                                       */
	uint8_t		wts[256];		/* scratch: Huffman weights                                             */
	uint8_t		ll_log, of_log, ml_log, huf_log;
	uint8_t		have_ll, have_of, have_ml, have_huf;
};

Z_HD int
z_highbit(uint32_t v)
{
	int			r = 0;

	while (v >>= 1)
		r++;
	return r;
}

/* ---- forward bit reading (FSE table descriptions) ---- */
Z_HD uint32_t
z_fwd_peek(const uint8_t *p, uint32_t len, uint32_t bitpos, int n)
{
	uint64_t	v = 0;
	const uint32_t b = bitpos >> 3;

	for (int i = 0; i < 5; i++)
		if (b + i < len)
			v |= (uint64_t) p[b + i] << (8 * i);
	return (uint32_t) ((v >> (bitpos & 7)) & ((1ull << n) - 1));
}

/* ---- backward bit streams: `pos` counts the unread bits; bits before the start of the stream read as zero ---- */
struct ZBits
{
	const uint8_t *p;
	int64_t		pos;
};

/* false = empty stream or last byte zero (no end mark) */
Z_HD bool
z_bits_init(ZBits &b, const uint8_t *p, uint32_t len)
{
	b.p = p;
	b.pos = 0;
	if (len == 0 || p[len - 1] == 0)
		return false;
	b.pos = (int64_t) (len - 1) * 8 + z_highbit(p[len - 1]);
	return true;
}

/* the n bits ending at `pos` (n <= 32), without consuming them */
Z_HD uint32_t
z_bits_peek_at(const ZBits &b, int64_t pos, int n)
{
	int64_t		lo = pos - n;
	uint64_t	v = 0;
	int			shift = 0;

	if (n == 0 || pos <= 0)
		return 0;
	if (lo < 0)
	{
		/* the part before the start of the stream is zero: the available bits are the high part */
		shift = (int) -lo;
		lo = 0;
	}
#ifdef __CUDA_ARCH__
	{
		/* two aligned words around the byte that holds bit `lo`: streams lie inside the file image on the device, which
		 * has readable bytes on both sides (block headers before, the inflate area / 16 spare bytes behind) */
		const uintptr_t a = (uintptr_t) (b.p + (lo >> 3));
		const uint64_t *w = (const uint64_t *) (a & ~(uintptr_t) 7);
		const uint32_t sh = (uint32_t) (a & 7) * 8u;
		uint64_t	x = w[0] >> sh;

		if (sh)
			x |= w[1] << (64u - sh);
		v = (x >> (lo & 7)) & ((1ull << (pos - lo)) - 1);
	}
#else
	{
		const int64_t byte = lo >> 3;
		const int	nb = (int) (((pos + 7) >> 3) - byte);	/* bytes that hold bits [lo, pos): at most 5 */

		for (int i = 0; i < nb; i++)
			v |= (uint64_t) b.p[byte + i] << (8 * i);
		v >>= (lo & 7);
		v &= (1ull << (pos - lo)) - 1;
	}
#endif
	return (uint32_t) (v << shift);
}

Z_HD uint32_t
z_bits_read(ZBits &b, int n)
{
	const uint32_t v = z_bits_peek_at(b, b.pos, n);

	b.pos -= n;
	return v;
}

/* ---- FSE ---- */
/* Normalised counts of an FSE table description (4.1.1).  Returns the bytes used, or -1. */
Z_HD int
z_read_ncount(const uint8_t *p, uint32_t len, int16_t *norm, int alphabet, int maxlog, int *logp)
{
	uint32_t	bp = 0;
	int			log,
				remaining,
				threshold,
				nbits,
				sym = 0;
	bool		prev0 = false;

	if (len == 0)
		return -1;
	log = (int) z_fwd_peek(p, len, bp, 4) + 5;
	bp += 4;
	if (log > maxlog)
		return -1;
	remaining = (1 << log) + 1;
	threshold = 1 << log;
	nbits = log + 1;
	while (remaining > 1 && sym < alphabet)
	{
		int			max,
					count;

		if (prev0)
		{
			int			n0 = sym;

			for (;;)
			{
				const uint32_t r = z_fwd_peek(p, len, bp, 2);

				bp += 2;
				n0 += (int) r;
				if (r != 3)
					break;
				if (bp > len * 8)
					return -1;
			}
			if (n0 > alphabet)
				return -1;
			while (sym < n0)
				norm[sym++] = 0;
			if (sym >= alphabet)
				break;
		}
		max = (2 * threshold - 1) - remaining;
		count = (int) z_fwd_peek(p, len, bp, nbits);
		if ((count & (threshold - 1)) < max)
		{
			count &= threshold - 1;
			bp += nbits - 1;
		}
		else
		{
			count &= 2 * threshold - 1;
			if (count >= threshold)
				count -= max;
			bp += nbits;
		}
		count--;
		remaining -= count < 0 ? -count : count;
		norm[sym++] = (int16_t) count;
		prev0 = count == 0;
		if (remaining < 1)
			return -1;
		while (remaining < threshold)
		{
			nbits--;
			threshold >>= 1;
		}
	}
	if (remaining != 1 || bp > len * 8)
		return -1;
	while (sym < alphabet)
		norm[sym++] = 0;
	*logp = log;
	return (int) ((bp + 7) >> 3);
}

/* decoding table from normalised counts (4.1.1, "from normalized distribution to decoding tables") */
Z_HD void
z_build_fse(ZFse *t, const int16_t *norm, int alphabet, int log, uint16_t *next)
{
	const int	size = 1 << log;
	const int	step = (size >> 1) + (size >> 3) + 3;
	const int	mask = size - 1;
	int			high = size - 1;
	int			pos = 0;

	for (int s = 0; s < alphabet; s++)
	{
		if (norm[s] == -1)
		{
			t[high--].sym = (uint8_t) s;
			next[s] = 1;
		}
		else
			next[s] = (uint16_t) norm[s];
	}
	for (int s = 0; s < alphabet; s++)
		for (int i = 0; i < norm[s]; i++)
		{
			t[pos].sym = (uint8_t) s;
			do
				pos = (pos + step) & mask;
			while (pos > high);
		}
	for (int u = 0; u < size; u++)
	{
		const int	s = t[u].sym;
		const uint32_t ns = next[s]++;
		const int	nb = log - z_highbit(ns);

		t[u].nbits = (uint8_t) nb;
		t[u].base = (uint16_t) ((ns << nb) - (uint32_t) size);
	}
}

Z_HD void
z_build_rle(ZFse *t, int sym)
{
	t[0].sym = (uint8_t) sym;
	t[0].nbits = 0;
	t[0].base = 0;
}

/* ---- Huffman ---- */
/*
 * Huffman_Tree_Description at p (4.2.1): weights, direct or FSE-compressed, then the decoding table.
 * Returns the bytes used, or -1.
 */
Z_HD int
z_huf_read_tree(const uint8_t *p, uint32_t len, ZTab &T)
{
	int			n = 0;			/* weights read */
	int			used;
	uint32_t	total = 0;
	int			maxbits;
	uint32_t	rest;

	if (len < 1)
		return -1;
	if (p[0] >= 128)
	{
		n = p[0] - 127;
		used = 1 + (n + 1) / 2;
		if ((uint32_t) used > len)
			return -1;
		for (int i = 0; i < n; i++)
			T.wts[i] = (i & 1) ? (p[1 + i / 2] & 15) : (p[1 + i / 2] >> 4);
	}
	else
	{
		const uint32_t csize = p[0];
		int			log = 0;
		int			hb;
		ZBits		bs;
		uint32_t	s1,
					s2;

		used = 1 + (int) csize;
		if (csize < 2 || (uint32_t) used > len)
			return -1;
		hb = z_read_ncount(p + 1, csize, T.norm, 13, 6, &log);
		if (hb < 0 || (uint32_t) hb >= csize)
			return -1;
		z_build_fse(T.wfse, T.norm, 13, log, T.next);
		if (!z_bits_init(bs, p + 1 + hb, csize - (uint32_t) hb))
			return -1;
		/* two interleaved states; the stream ends when an update runs out of bits (4.2.1.2) */
		s1 = z_bits_read(bs, log);
		s2 = z_bits_read(bs, log);
		if (bs.pos < 0)
			return -1;
		for (;;)
		{
			if (n >= 254)
				return -1;
			T.wts[n++] = T.wfse[s1].sym;
			s1 = T.wfse[s1].base + z_bits_read(bs, T.wfse[s1].nbits);
			if (bs.pos < 0)
			{
				T.wts[n++] = T.wfse[s2].sym;
				break;
			}
			if (n >= 254)
				return -1;
			T.wts[n++] = T.wfse[s2].sym;
			s2 = T.wfse[s2].base + z_bits_read(bs, T.wfse[s2].nbits);
			if (bs.pos < 0)
			{
				T.wts[n++] = T.wfse[s1].sym;
				break;
			}
		}
	}
	/* the last weight is what completes the sum to a power of two */
	for (int i = 0; i < n; i++)
	{
		if (T.wts[i] > Z_HUF_MAXLOG)
			return -1;
		if (T.wts[i])
			total += 1u << (T.wts[i] - 1);
	}
	if (total == 0)
		return -1;
	maxbits = z_highbit(total) + 1;
	if (maxbits > Z_HUF_MAXLOG)
		return -1;
	rest = (1u << maxbits) - total;
	if (rest & (rest - 1))
		return -1;				/* not a power of two */
	T.wts[n] = (uint8_t) (z_highbit(rest) + 1);
	n++;
	/* table: lowest weights (longest codes) first, symbols in order within a weight */
	{
		uint32_t	pos = 0;

		for (int w = 1; w <= maxbits; w++)
			for (int s = 0; s < n; s++)
				if (T.wts[s] == w)
				{
					const uint32_t cnt = 1u << (w - 1);
					const uint16_t e = (uint16_t) (s | ((maxbits + 1 - w) << 8));

					for (uint32_t i = 0; i < cnt; i++)
						T.huf[pos + i] = e;
					pos += cnt;
				}
		if (pos != (1u << maxbits))
			return -1;
	}
	T.huf_log = (uint8_t) maxbits;
	T.have_huf = 1;
	return used;
}

/* one Huffman-coded stream of `count` symbols into out; false = the stream does not end exactly where it should */
Z_HD bool
z_huf_stream(const ZTab &T, const uint8_t *p, uint32_t len, uint8_t *out, uint32_t count)
{
	ZBits		bs;
	const int	log = T.huf_log;

	if (!z_bits_init(bs, p, len))
		return false;
	for (uint32_t i = 0; i < count; i++)
	{
		const uint16_t e = T.huf[z_bits_peek_at(bs, bs.pos, log)];

		out[i] = (uint8_t) e;
		bs.pos -= e >> 8;
		if (bs.pos < 0)
			return false;
	}
	return bs.pos == 0;
}

/* ---- literals section ---- */
struct ZLit
{
	int			type;			/* 0 raw, 1 RLE, 2 compressed, 3 treeless                               */
	uint32_t	regen;			/* literals in the block                                                */
	uint32_t	csize;			/* compressed: bytes of tree description + streams                      */
	int			streams;		/* 1 or 4                                                               */
	uint32_t	hdr;			/* header bytes                                                         */
};

/* Literals_Section_Header (3.1.1.3.1.1); false = malformed */
Z_HD bool
z_lit_header(const uint8_t *p, uint32_t len, ZLit &L)
{
	uint32_t	h;
	int			fmt;

	if (len < 1)
		return false;
	L.type = p[0] & 3;
	fmt = (p[0] >> 2) & 3;
	L.streams = 1;
	L.csize = 0;
	if (L.type < 2)
	{
		if ((fmt & 1) == 0)
		{
			L.hdr = 1;
			L.regen = p[0] >> 3;
		}
		else if (fmt == 1)
		{
			if (len < 2)
				return false;
			L.hdr = 2;
			L.regen = ((uint32_t) p[0] | ((uint32_t) p[1] << 8)) >> 4;
		}
		else
		{
			if (len < 3)
				return false;
			L.hdr = 3;
			L.regen = ((uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16)) >> 4;
		}
		return true;
	}
	if (len < 3)
		return false;
	h = (uint32_t) p[0] | ((uint32_t) p[1] << 8) | ((uint32_t) p[2] << 16);
	if (fmt < 2)
	{
		L.hdr = 3;
		L.streams = fmt == 0 ? 1 : 4;
		L.regen = (h >> 4) & 0x3FF;
		L.csize = (h >> 14) & 0x3FF;
	}
	else if (fmt == 2)
	{
		if (len < 4)
			return false;
		h |= (uint32_t) p[3] << 24;
		L.hdr = 4;
		L.streams = 4;
		L.regen = (h >> 4) & 0x3FFF;
		L.csize = h >> 18;
	}
	else
	{
		uint64_t	h5;

		if (len < 5)
			return false;
		h5 = (uint64_t) h | ((uint64_t) p[3] << 24) | ((uint64_t) p[4] << 32);
		L.hdr = 5;
		L.streams = 4;
		L.regen = (uint32_t) ((h5 >> 4) & 0x3FFFF);
		L.csize = (uint32_t) (h5 >> 22);
	}
	return true;
}

/* where the (up to) four Huffman streams of a compressed literals section lie, and how many symbols each holds */
struct ZLitStreams
{
	uint32_t	off[4];
	uint32_t	len[4];
	uint32_t	count[4];
	uint32_t	outoff[4];
};

Z_HD bool
z_lit_streams(const ZLit &L, const uint8_t *p /* after the tree */ , uint32_t len, ZLitStreams &S)
{
	if (L.streams == 1)
	{
		S.off[0] = 0;
		S.len[0] = len;
		S.count[0] = L.regen;
		S.outoff[0] = 0;
		for (int i = 1; i < 4; i++)
			S.off[i] = S.len[i] = S.count[i] = S.outoff[i] = 0;
		return true;
	}
	if (len < 6)
		return false;
	{
		const uint32_t s1 = (uint32_t) p[0] | ((uint32_t) p[1] << 8);
		const uint32_t s2 = (uint32_t) p[2] | ((uint32_t) p[3] << 8);
		const uint32_t s3 = (uint32_t) p[4] | ((uint32_t) p[5] << 8);
		const uint32_t per = (L.regen + 3) / 4;

		if ((uint64_t) 6 + s1 + s2 + s3 > len || per * 3 > L.regen)
			return false;
		S.off[0] = 6;
		S.len[0] = s1;
		S.off[1] = 6 + s1;
		S.len[1] = s2;
		S.off[2] = 6 + s1 + s2;
		S.len[2] = s3;
		S.off[3] = 6 + s1 + s2 + s3;
		S.len[3] = len - S.off[3];
		for (int i = 0; i < 4; i++)
		{
			S.count[i] = i < 3 ? per : L.regen - 3 * per;
			S.outoff[i] = (uint32_t) i * per;
		}
	}
	return true;
}

/* ---- sequences section ---- */
struct ZSeqState
{
	ZBits		bs;
	uint32_t	ll_state, of_state, ml_state;
	uint32_t	nseq;
	uint32_t	done;
	uint32_t	rep[3];			/* repeat offsets: kept across the blocks of a frame                    */
};

/* one of the three tables (3.1.1.3.2.1): mode 0 predefined, 1 RLE, 2 FSE description, 3 repeat.  Returns bytes used or -1 */
Z_HD int
z_seq_table(int mode, const uint8_t *p, uint32_t len, ZFse *t, uint8_t *logp, uint8_t *have, const int16_t *defnorm, int deflog,
			int alphabet, int maxlog, ZTab &T)
{
	if (mode == 0)
	{
		for (int i = 0; i < alphabet; i++)
			T.norm[i] = defnorm[i];
		z_build_fse(t, T.norm, alphabet, deflog, T.next);
		*logp = (uint8_t) deflog;
		*have = 1;
		return 0;
	}
	if (mode == 1)
	{
		if (len < 1 || p[0] >= alphabet)
			return -1;
		z_build_rle(t, p[0]);
		*logp = 0;
		*have = 1;
		return 1;
	}
	if (mode == 2)
	{
		int			log = 0;
		const int	used = z_read_ncount(p, len, T.norm, alphabet, maxlog, &log);

		if (used < 0)
			return -1;
		z_build_fse(t, T.norm, alphabet, log, T.next);
		*logp = (uint8_t) log;
		*have = 1;
		return used;
	}
	return *have ? 0 : -1;
}

/*
 * Sequences_Section_Header + tables + bitstream initialisation.  p / len = the rest of the block after the literals.
 * Returns false on malformed input.  nseq == 0: no bitstream.
 */
Z_HD bool
z_seq_begin(const uint8_t *p, uint32_t len, ZTab &T, ZSeqState &S)
{
	uint32_t	pos = 0;
	int			modes,
				used;

	S.done = 0;
	if (len < 1)
		return false;
	if (p[0] == 0)
	{
		S.nseq = 0;
		return len == 1;
	}
	if (p[0] < 128)
	{
		S.nseq = p[0];
		pos = 1;
	}
	else if (p[0] < 255)
	{
		if (len < 2)
			return false;
		S.nseq = (((uint32_t) p[0] - 128u) << 8) + p[1];
		pos = 2;
	}
	else
	{
		if (len < 3)
			return false;
		S.nseq = (uint32_t) p[1] + ((uint32_t) p[2] << 8) + 0x7F00u;
		pos = 3;
	}
	if (pos >= len)
		return false;
	modes = p[pos++];
	if (modes & 3)
		return false;			/* reserved bits */
	used = z_seq_table((modes >> 6) & 3, p + pos, len - pos, T.ll, &T.ll_log, &T.have_ll, Z_T(z_ll_default), 6, 36, Z_LL_MAXLOG, T);
	if (used < 0)
		return false;
	pos += (uint32_t) used;
	used = z_seq_table((modes >> 4) & 3, p + pos, len - pos, T.of, &T.of_log, &T.have_of, Z_T(z_of_default), 5, 29, Z_OF_MAXLOG, T);
	if (used < 0)
		return false;
	pos += (uint32_t) used;
	used = z_seq_table((modes >> 2) & 3, p + pos, len - pos, T.ml, &T.ml_log, &T.have_ml, Z_T(z_ml_default), 6, 53, Z_ML_MAXLOG, T);
	if (used < 0)
		return false;
	pos += (uint32_t) used;
	if (pos >= len || !z_bits_init(S.bs, p + pos, len - pos))
		return false;
	S.ll_state = z_bits_read(S.bs, T.ll_log);
	S.of_state = z_bits_read(S.bs, T.of_log);
	S.ml_state = z_bits_read(S.bs, T.ml_log);
	return S.bs.pos >= 0;
}

/* the next sequence: literal length, match length, offset (repeat offsets resolved).  false = malformed */
Z_HD bool
z_seq_next(const ZTab &T, ZSeqState &S, uint32_t *llp, uint32_t *mlp, uint32_t *offp)
{
	const ZFse	le = T.ll[S.ll_state];
	const ZFse	oe = T.of[S.of_state];
	const ZFse	me = T.ml[S.ml_state];
	uint32_t	ofv,
				ml,
				ll,
				off;

	if (oe.sym > 31 || le.sym > 35 || me.sym > 52)
		return false;
	ofv = (1u << oe.sym) + z_bits_read(S.bs, oe.sym);
	ml = Z_T(z_ml_base)[me.sym] + z_bits_read(S.bs, Z_T(z_ml_bits)[me.sym]);
	ll = Z_T(z_ll_base)[le.sym] + z_bits_read(S.bs, Z_T(z_ll_bits)[le.sym]);
	if (ofv > 3)
	{
		off = ofv - 3;
		S.rep[2] = S.rep[1];
		S.rep[1] = S.rep[0];
		S.rep[0] = off;
	}
	else
	{
		uint32_t	idx = ofv - 1;	/* 0, 1, 2 */

		if (ll == 0)
			idx++;
		if (idx == 0)
			off = S.rep[0];
		else
		{
			off = idx < 3 ? S.rep[idx] : S.rep[0] - 1;
			if (off == 0)
				return false;
			if (idx > 1)
				S.rep[2] = S.rep[1];
			S.rep[1] = S.rep[0];
			S.rep[0] = off;
		}
	}
	S.done++;
	if (S.done < S.nseq)
	{
		S.ll_state = le.base + z_bits_read(S.bs, le.nbits);
		S.ml_state = me.base + z_bits_read(S.bs, me.nbits);
		S.of_state = oe.base + z_bits_read(S.bs, oe.nbits);
	}
	if (S.bs.pos < 0)
		return false;
	*llp = ll;
	*mlp = ml;
	*offp = off;
	return true;
}

/* ---- frame ---- */
struct ZFrame
{
	uint32_t	hdr;			/* bytes before the first block                                         */
	uint64_t	content_size;	/* Frame_Content_Size, or ~0 when the header does not carry it          */
	int			checksum;		/* Content_Checksum_flag                                                */
};

Z_HD bool
z_frame_header(const uint8_t *p, uint32_t len, ZFrame &F)
{
	uint32_t	pos = 5;
	int			fhd,
				fcs,
				single,
				did;

	if (len < 6 || p[0] != 0x28 || p[1] != 0xB5 || p[2] != 0x2F || p[3] != 0xFD)
		return false;
	fhd = p[4];
	fcs = fhd >> 6;
	single = (fhd >> 5) & 1;
	did = fhd & 3;
	if (fhd & 0x08)
		return false;			/* reserved bit */
	F.checksum = (fhd >> 2) & 1;
	if (did)
		return false;			/* dictionaries: not used by the reference */
	if (!single)
		pos++;					/* Window_Descriptor: offsets are checked against the output produced so far */
	F.content_size = ~0ull;
	{
		const int	n = fcs == 0 ? (single ? 1 : 0) : fcs == 1 ? 2 : fcs == 2 ? 4 : 8;
		uint64_t	v = 0;

		if (pos + (uint32_t) n > len)
			return false;
		for (int i = 0; i < n; i++)
			v |= (uint64_t) p[pos + i] << (8 * i);
		if (n == 2)
			v += 256;
		if (n)
			F.content_size = v;
		pos += (uint32_t) n;
	}
	F.hdr = pos;
	return true;
}

/* Block_Header (3.1.1.2): 3 bytes */
Z_HD bool
z_block_header(const uint8_t *p, uint32_t len, uint32_t pos, int *last, int *type, uint32_t *size)
{
	uint32_t	h;

	if (pos + 3 > len)
		return false;
	h = (uint32_t) p[pos] | ((uint32_t) p[pos + 1] << 8) | ((uint32_t) p[pos + 2] << 16);
	*last = (int) (h & 1);
	*type = (int) ((h >> 1) & 3);
	*size = h >> 3;
	return *type != 3 && (*type == 1 || *size <= Z_BLOCK_MAX);
}

/* XXH64 of the content, seed 0: the low 32 bits are the frame's Content_Checksum */
Z_HD uint64_t
z_xxh_rotl(uint64_t v, int r)
{
	return (v << r) | (v >> (64 - r));
}

Z_HD uint64_t
z_xxh_read64(const uint8_t *p)
{
	uint64_t	v = 0;

	for (int i = 0; i < 8; i++)
		v |= (uint64_t) p[i] << (8 * i);
	return v;
}

Z_HD uint64_t
z_xxh_round(uint64_t acc, uint64_t in)
{
	acc += in * 0xC2B2AE3D27D4EB4Full;
	acc = z_xxh_rotl(acc, 31);
	return acc * 0x9E3779B185EBCA87ull;
}

Z_HD uint64_t
z_xxh_merge(uint64_t acc, uint64_t v)
{
	acc ^= z_xxh_round(0, v);
	return acc * 0x9E3779B185EBCA87ull + 0x85EBCA77C2B2AE63ull;
}

Z_HD uint64_t
z_xxh64(const uint8_t *p, uint64_t len)
{
	const uint64_t P1 = 0x9E3779B185EBCA87ull, P2 = 0xC2B2AE3D27D4EB4Full, P3 = 0x165667B19E3779F9ull, P4 = 0x85EBCA77C2B2AE63ull,
				P5 = 0x27D4EB2F165667C5ull;
	uint64_t	h;
	uint64_t	i = 0;

	if (len >= 32)
	{
		uint64_t	v1 = P1 + P2, v2 = P2, v3 = 0, v4 = 0 - P1;

		for (; i + 32 <= len; i += 32)
		{
			v1 = z_xxh_round(v1, z_xxh_read64(p + i));
			v2 = z_xxh_round(v2, z_xxh_read64(p + i + 8));
			v3 = z_xxh_round(v3, z_xxh_read64(p + i + 16));
			v4 = z_xxh_round(v4, z_xxh_read64(p + i + 24));
		}
		h = z_xxh_rotl(v1, 1) + z_xxh_rotl(v2, 7) + z_xxh_rotl(v3, 12) + z_xxh_rotl(v4, 18);
		h = z_xxh_merge(h, v1);
		h = z_xxh_merge(h, v2);
		h = z_xxh_merge(h, v3);
		h = z_xxh_merge(h, v4);
	}
	else
		h = P5;
	h += len;
	for (; i + 8 <= len; i += 8)
	{
		h ^= z_xxh_round(0, z_xxh_read64(p + i));
		h = z_xxh_rotl(h, 27) * P1 + P4;
	}
	if (i + 4 <= len)
	{
		h ^= ((uint64_t) p[i] | ((uint64_t) p[i + 1] << 8) | ((uint64_t) p[i + 2] << 16) | ((uint64_t) p[i + 3] << 24)) * P1;
		h = z_xxh_rotl(h, 23) * P2 + P3;
		i += 4;
	}
	for (; i < len; i++)
	{
		h ^= p[i] * P5;
		h = z_xxh_rotl(h, 11) * P1;
	}
	h ^= h >> 33;
	h *= P2;
	h ^= h >> 29;
	h *= P3;
	h ^= h >> 32;
	return h;
}

#endif							/* CB_ZSTD_DEC_CUH */
