/* Host build of the product's serial DEFLATE decoder (cloudberry_b200/csrc/inflate.cuh, the __host__ __device__ part
 * that lane 0 of each warp runs in k_aocs_inflate), with the queue applied serially.  Test harness only: compiled by
 * tests/test_inflate_host.py with g++ and compared against the system zlib. */
#include <string.h>
#include "../../cloudberry_b200/csrc/inflate.cuh"

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
		const int	rc = infl_step(s, T, q, &n);

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
