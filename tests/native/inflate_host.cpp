/* Host build of the product's serial DEFLATE decoder (cloudberry_b200/csrc/inflate.cuh, the __host__ __device__ part
 * that lane 0 of each warp runs in k_aocs_inflate), with the queue applied serially.  Test harness only: compiled by
 * tests/test_inflate_host.py with g++ and compared against the system zlib. */
#include <string.h>
#include "../../cloudberry_b200/csrc/inflate.cuh"

/* The batch schedule of infl_apply_warp (aocs.cu) replayed on the host: positions from a prefix sum; literals and
 * matches whose source ends before the batch go first -- here in REVERSE lane order, the most hostile interleaving a
 * warp could produce -- then the matches that read the batch's own output, in order, each in 32-wide strides.  The
 * result must be what the serial replay gives. */
static bool
apply_like_a_warp(uint8_t *out, uint32_t &outpos, uint32_t cap, const uint32_t *q, int n)
{
	uint32_t	start[INFL_QN], len[INFL_QN], dist[INFL_QN];
	bool		dep[INFL_QN];
	uint32_t	p = outpos;

	for (int i = 0; i < n; i++)
	{
		len[i] = (q[i] & INFL_LIT) ? 1u : (q[i] & 511u);
		dist[i] = (q[i] >> 9) & 0xFFFFu;
		start[i] = p;
		p += len[i];
		if (!(q[i] & INFL_LIT) && dist[i] > start[i])
			return false;
	}
	if (p > cap)
		return false;
	for (int i = n - 1; i >= 0; i--)
	{
		dep[i] = false;
		if (q[i] & INFL_LIT)
			out[start[i]] = (uint8_t) q[i];
		else if (start[i] - dist[i] + len[i] <= outpos)
			for (uint32_t j = 0; j < len[i]; j++)
				out[start[i] + j] = out[start[i] - dist[i] + j];
		else
			dep[i] = true;
	}
	for (int i = 0; i < n; i++)
		if (dep[i])
			for (uint32_t base = 0; base < len[i]; base += 32)
			{
				uint8_t		tmp[32];
				uint32_t	m = len[i] - base < 32 ? len[i] - base : 32;

				/* one stride: all 32 loads happen before any of the 32 stores becomes visible, or after: both orders
				 * must agree, so read everything first */
				for (uint32_t l = 0; l < m; l++)
				{
					const uint32_t j = base + l;

					tmp[l] = out[start[i] - dist[i] + (dist[i] >= len[i] ? j : j % dist[i])];
				}
				for (uint32_t l = 0; l < m; l++)
					out[start[i] + base + l] = tmp[l];
			}
	outpos = p;
	return true;
}

extern "C" long long
infl_host_zlib_warp(const uint8_t *z, uint32_t zl, uint8_t *out, uint32_t cap, uint32_t *adler_stored)
{
	static thread_local InflTables T;
	InflState	s;
	uint32_t	q[INFL_QN];
	uint32_t	pos = 0;
	int			n;

	if (!infl_zlib_header_ok(z, zl))
		return -1;
	infl_init(s, z, zl, 2);
	for (;;)
	{
		const int	rc = infl_step<1>(s, infl_view(T), q, &n);

		if (rc == INFL_ERROR)
			return -2;
		if (n && !apply_like_a_warp(out, pos, cap, q, n))
			return -4;
		if (rc == INFL_STORED)
		{
			if (pos + s.stored_len > cap)
				return -5;
			memcpy(out + pos, z + s.stored_src, s.stored_len);
			pos += s.stored_len;
		}
		if (rc == INFL_DONE)
			break;
	}
	const uint32_t c = infl_consumed(s);

	if (c + 4 > zl)
		return -6;
	*adler_stored = ((uint32_t) z[c] << 24) | ((uint32_t) z[c + 1] << 16) | ((uint32_t) z[c + 2] << 8) | z[c + 3];
	return pos;
}

extern "C" long long
infl_host_zlib(const uint8_t *z, uint32_t zl, uint8_t *out, uint32_t cap, uint32_t *adler_stored)
{
	static thread_local InflTables T;
	InflState	s;
	uint32_t	q[INFL_QN];
	uint32_t	pos = 0;
	int			n;

	if (!infl_zlib_header_ok(z, zl))
		return -1;
	infl_init(s, z, zl, 2);
	for (;;)
	{
		const int	rc = infl_step<1>(s, infl_view(T), q, &n);

		if (rc == INFL_ERROR)
			return -2;
		for (int i = 0; i < n; i++)
		{
			const uint32_t e = q[i];

			if (e & INFL_LIT)
			{
				if (pos >= cap)
					return -3;
				out[pos++] = (uint8_t) e;
			}
			else
			{
				const uint32_t l = e & 511u, d = (e >> 9) & 0xFFFFu;

				if (d > pos || pos + l > cap)
					return -4;
				for (uint32_t j = 0; j < l; j++, pos++)
					out[pos] = out[pos - d];
			}
		}
		if (rc == INFL_STORED)
		{
			if (pos + s.stored_len > cap)
				return -5;
			memcpy(out + pos, z + s.stored_src, s.stored_len);
			pos += s.stored_len;
		}
		if (rc == INFL_DONE)
			break;
	}
	const uint32_t c = infl_consumed(s);

	if (c + 4 > zl)
		return -6;
	*adler_stored = ((uint32_t) z[c] << 24) | ((uint32_t) z[c + 1] << 16) | ((uint32_t) z[c + 2] << 8) | z[c + 3];
	return pos;
}
