/* Host build of the hash functions the kernels inline (cloudberry_b200/csrc/common.cuh, the __host__ __device__ ones: the same
 * source text nvcc compiles into every join / aggregate / Motion kernel).  Test harness only: compiled by
 * tests/test_device_hash_source.py with nvcc (host side of the translation unit) and compared with the reference's own
 * hashfunc.c / varchar.c / cdbhash.c / hashfn.h as compiled into oracle/_ref/libexec_ref.so. */
#include "../../cloudberry_b200/csrc/common.cuh"

extern "C" {
uint32_t	hh_hash_uint32(uint32_t k) { return pg_hash_uint32(k); }
uint32_t	hh_hashint8(int64_t v) { return pg_hashint8(v); }
uint32_t	hh_hashfloat8(uint64_t bits) { return pg_hashfloat8(bits); }
uint32_t	hh_hash_bpchar1(uint8_t ch) { return pg_hash_bpchar1(ch); }
uint32_t	hh_hash_bytes(const unsigned char *k, int len) { return pg_hash_bytes_host(k, len); }
uint32_t	hh_murmurhash32(uint32_t h) { return pg_murmurhash32(h); }
uint32_t	hh_hash_combine(uint32_t acc, uint32_t h, int isnull) { return pg_hash_combine(acc, h, isnull != 0); }
int32_t		hh_jump_consistent_hash(uint64_t key, int32_t nseg) { return pg_jump_consistent_hash(key, nseg); }
}
