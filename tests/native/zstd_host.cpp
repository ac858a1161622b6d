/* Host build of the product's serial Zstandard decoder (cloudberry_b200/csrc/zstd_dec.cuh), with the frame / block loop
 * and the sequence execution done serially.  Test harness only: compiled by tests/test_zstd_host.py with g++ and
 * compared against a real libzstd (pyarrow's). */
#include <string.h>
#include <stdlib.h>
#include "../../cloudberry_b200/csrc/zstd_dec.cuh"

extern "C" unsigned long long
zstd_host_xxh64(const uint8_t *p, unsigned long long n)
{
	return z_xxh64(p, n);
}

/* which format features the streams fed so far exercised (so the test can insist its corpus covers them):
 * 0-2 block raw/rle/compressed, 3-6 literals raw/rle/compressed/treeless, 7 four streams, 8 one stream, 9 direct weights,
 * 10 FSE weights, 11-14 LL mode 0-3, 15-18 OF mode, 19-22 ML mode, 23 frames with more than one block, 24 nseq == 0,
 * 25 long nseq header (>= 128), 26 repeat offset used with ll == 0 */
static unsigned long long z_stats[32];

extern "C" unsigned long long
zstd_host_stat(int i)
{
	return z_stats[i];
}

/* The batch schedule of zstd_exec_warp (aocs.cu) replayed on the host, phase 1 in REVERSE lane order: literal runs and
 * matches whose source ends before the batch first, then the matches that read the batch's own output (or are long), in
 * order.  Must give what the serial execution gives. */
static bool
exec_like_a_warp(uint8_t *out, uint32_t &op, uint32_t cap, const uint8_t *lit, uint32_t &lp, uint32_t regen, const uint32_t *qll,
				 const uint32_t *qml, const uint32_t *qoff, int n)
{
	uint32_t	start[Z_SEQ_QN], lstart[Z_SEQ_QN];
	bool		dep[Z_SEQ_QN];
	uint32_t	o = op, l = lp;

	for (int i = 0; i < n; i++)
	{
		start[i] = o;
		lstart[i] = l;
		o += qll[i] + qml[i];
		l += qll[i];
		if (qoff[i] == 0 || qoff[i] > start[i] + qll[i])
			return false;
	}
	if (l > regen || o > cap)
		return false;
	for (int i = n - 1; i >= 0; i--)
	{
		const uint32_t ms = start[i] + qll[i];

		memcpy(out + start[i], lit + lstart[i], qll[i]);
		dep[i] = false;
		if (ms - qoff[i] + qml[i] <= op && qml[i] <= 64)
			for (uint32_t j = 0; j < qml[i]; j++)
				out[ms + j] = out[ms - qoff[i] + j];
		else
			dep[i] = true;
	}
	for (int i = 0; i < n; i++)
		if (dep[i])
		{
			const uint32_t ms = start[i] + qll[i], d = qoff[i], len = qml[i];

			for (uint32_t j = 0; j < len; j++)
				out[ms + j] = out[ms - d + (d >= len ? j : j % d)];
		}
	op = o;
	lp = l;
	return true;
}

static int	z_warp_mode;

extern "C" void
zstd_host_set_warp_mode(int on)
{
	z_warp_mode = on;
}

/* returns bytes produced, or a negative stage code */
extern "C" long long
zstd_host_decompress(const uint8_t *z, uint32_t zl, uint8_t *out, uint32_t cap)
{
	static thread_local ZTab T;
	static thread_local uint8_t litbuf[Z_BLOCK_MAX + 32];
	ZFrame		F;
	ZSeqState	S;
	uint32_t	pos,
				op = 0;
	int			last = 0;

	memset(&T, 0, sizeof(T));
	if (!z_frame_header(z, zl, F))
		return -1;
	pos = F.hdr;
	S.rep[0] = 1;
	S.rep[1] = 4;
	S.rep[2] = 8;
	while (!last)
	{
		int			type;
		uint32_t	size;

		if (!z_block_header(z, zl, pos, &last, &type, &size))
			return -2;
		pos += 3;
		z_stats[type]++;
		if (!last)
			z_stats[23]++;
		if (type == 0)
		{
			if (pos + size > zl || op + size > cap)
				return -3;
			memcpy(out + op, z + pos, size);
			op += size;
			pos += size;
			continue;
		}
		if (type == 1)
		{
			if (pos + 1 > zl || op + size > cap)
				return -4;
			memset(out + op, z[pos], size);
			op += size;
			pos += 1;
			continue;
		}
		{
			const uint8_t *b = z + pos;
			ZLit		L;
			const uint8_t *lit;
			uint32_t	lp = 0,
						after;

			if (pos + size > zl || !z_lit_header(b, size, L) || L.regen > Z_BLOCK_MAX)
				return -5;
			z_stats[3 + L.type]++;
			if (L.type >= 2)
				z_stats[L.streams == 4 ? 7 : 8]++;
			if (L.type == 0)
			{
				if (L.hdr + L.regen > size)
					return -6;
				lit = b + L.hdr;
				after = L.hdr + L.regen;
			}
			else if (L.type == 1)
			{
				if (L.hdr + 1 > size)
					return -7;
				memset(litbuf, b[L.hdr], L.regen);
				lit = litbuf;
				after = L.hdr + 1;
			}
			else
			{
				uint32_t	tree = 0;
				ZLitStreams St;

				if (L.hdr + L.csize > size)
					return -8;
				if (L.type == 2)
				{
					const int	used = z_huf_read_tree(b + L.hdr, L.csize, T);

					z_stats[b[L.hdr] >= 128 ? 9 : 10]++;
					if (used < 0)
						return -9;
					tree = (uint32_t) used;
				}
				else if (!T.have_huf)
					return -10;
				if (!z_lit_streams(L, b + L.hdr + tree, L.csize - tree, St))
					return -11;
				for (int i = 0; i < L.streams; i++)
					if (!z_huf_stream(T, b + L.hdr + tree + St.off[i], St.len[i], litbuf + St.outoff[i], St.count[i]))
						return -12;
				lit = litbuf;
				after = L.hdr + L.csize;
			}
			if (!z_seq_begin(b + after, size - after, T, S))
				return -13;
			if (S.nseq == 0)
				z_stats[24]++;
			else
			{
				const int	m = b[after + (b[after] < 128 ? 1 : b[after] < 255 ? 2 : 3)];

				z_stats[11 + ((m >> 6) & 3)]++;
				z_stats[15 + ((m >> 4) & 3)]++;
				z_stats[19 + ((m >> 2) & 3)]++;
				if (b[after] >= 128)
					z_stats[25]++;
			}
			if (z_warp_mode)
			{
				for (uint32_t done = 0; done < S.nseq; done += Z_SEQ_QN)
				{
					uint32_t	qll[Z_SEQ_QN], qml[Z_SEQ_QN], qoff[Z_SEQ_QN];
					const int	n = (int) (S.nseq - done < Z_SEQ_QN ? S.nseq - done : Z_SEQ_QN);

					for (int k = 0; k < n; k++)
						if (!z_seq_next(T, S, &qll[k], &qml[k], &qoff[k]))
							return -14;
					if (!exec_like_a_warp(out, op, cap, lit, lp, L.regen, qll, qml, qoff, n))
						return -15;
				}
			}
			else
				for (uint32_t i = 0; i < S.nseq; i++)
				{
					uint32_t	ll, ml, off;

					if (!z_seq_next(T, S, &ll, &ml, &off))
						return -14;
					if (lp + ll > L.regen || op + ll + ml > cap || off > op + ll)
						return -15;
					memcpy(out + op, lit + lp, ll);
					op += ll;
					lp += ll;
					for (uint32_t j = 0; j < ml; j++, op++)
						out[op] = out[op - off];
				}
			if (S.nseq && S.bs.pos != 0)
				return -16;
			if (op + (L.regen - lp) > cap)
				return -17;
			memcpy(out + op, lit + lp, L.regen - lp);
			op += L.regen - lp;
			pos += size;
		}
	}
	if (F.content_size != ~0ull && F.content_size != op)
		return -18;
	if (F.checksum)
	{
		if (pos + 4 > zl)
			return -19;
		const uint32_t want = (uint32_t) z[pos] | ((uint32_t) z[pos + 1] << 8) | ((uint32_t) z[pos + 2] << 16) | ((uint32_t) z[pos + 3] << 24);

		if ((uint32_t) z_xxh64(out, op) != want)
			return -20;
	}
	return op;
}
