/*
 * fake_cudart.c - a CUDA runtime that computes NOTHING, for unit tests of the HOST code around the kernels.
 *
 * TEST INFRASTRUCTURE ONLY (tests/test_host_logic_fake_runtime.py builds it into a temporary directory and LD_PRELOADs it
 * into a subprocess; nothing in the product, bench.py or smoke() ever loads it).  It is not a CPU fallback: "device" memory
 * is zero-filled host memory and every kernel launch is a no-op, so no query result comes out of it.  What it makes testable
 * without a GPU is everything the host does on the way to and from the kernels: plan validation and the refusal messages
 * of cb_exec.c, pipeline program emission, the specialised-kernel matcher, launch bookkeeping, result read-back paths,
 * clean-up - under the sanitizers if wanted.  The runtime entry points below are exactly the ones libcbgpu.so imports.
 */
#include <cuda_runtime_api.h>
#include <stdlib.h>
#include <string.h>

static long	launches;

long		fake_cudart_launches(void) { return launches; }

cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { (void) d; return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "fake runtime error"; }

cudaError_t
cudaGetDeviceProperties_v2(struct cudaDeviceProp *p, int d)
{
	(void) d;
	memset(p, 0, sizeof(*p));
	strcpy(p->name, "fake (no kernels run)");
	p->major = 10;
	p->minor = 0;
	p->multiProcessorCount = 148;
	p->totalGlobalMem = (size_t) 180 << 30;
	p->sharedMemPerBlock = 48 << 10;
	p->sharedMemPerBlockOptin = 227 << 10;
	p->sharedMemPerMultiprocessor = 228 << 10;
	p->regsPerMultiprocessor = 65536;
	p->maxThreadsPerBlock = 1024;
	p->maxThreadsPerMultiProcessor = 2048;
	p->warpSize = 32;
	p->l2CacheSize = 126 << 20;
	return cudaSuccess;
}

cudaError_t cudaMemGetInfo(size_t *fr, size_t *tot) { *fr = (size_t) 170 << 30; *tot = (size_t) 180 << 30; return cudaSuccess; }

/* fault injection: the n-th device allocation from now fails (0 = off) - the host's out-of-memory paths */
static long	fail_alloc_in;

void		fake_cudart_fail_alloc_in(long n) { fail_alloc_in = n; }

static cudaError_t
zalloc_dev(void **p, size_t n)
{
	if (fail_alloc_in > 0 && --fail_alloc_in == 0)
	{
		*p = NULL;
		return cudaErrorMemoryAllocation;
	}
	*p = calloc(1, n ? n : 1);
	return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}

static cudaError_t
zalloc(void **p, size_t n)
{
	*p = calloc(1, n ? n : 1);
	return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}

cudaError_t cudaMalloc(void **p, size_t n) { return zalloc_dev(p, n); }
cudaError_t cudaMallocAsync(void **p, size_t n, cudaStream_t s) { (void) s; return zalloc_dev(p, n); }
cudaError_t cudaMallocHost(void **p, size_t n) { return zalloc(p, n); }
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned int f) { (void) f; return zalloc(p, n); }
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaFreeAsync(void *p, cudaStream_t s) { (void) s; free(p); return cudaSuccess; }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }

/* fault injection: the n-th 4-byte device-to-host copy from now delivers `value` instead of what the "device" holds - a status
 * word reporting an error, or a count that makes no sense (0 = off) */
static long	poison_in;
static int	poison_value;

void		fake_cudart_poison_int_d2h(long n, int value) { poison_in = n; poison_value = value; }

cudaError_t
cudaMemcpyAsync(void *dst, const void *src, size_t n, enum cudaMemcpyKind k, cudaStream_t s)
{
	(void) s;
	memmove(dst, src, n);
	if (k == cudaMemcpyDeviceToHost && n == sizeof(int) && poison_in > 0 && --poison_in == 0)
		memcpy(dst, &poison_value, sizeof(int));
	return cudaSuccess;
}

cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t s) { (void) s; memset(p, v, n); return cudaSuccess; }

cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t *pool, int d) { (void) d; *pool = (cudaMemPool_t) (void *) &launches; return cudaSuccess; }
cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t pool, enum cudaMemPoolAttr a, void *v) { (void) pool; (void) a; (void) v; return cudaSuccess; }

cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned int f) { (void) f; *s = (cudaStream_t) calloc(1, 8); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t s) { (void) s; return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t) calloc(1, 8); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) { (void) e; (void) s; return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t e) { (void) e; return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { (void) a; (void) b; *ms = 0.001f; return cudaSuccess; }

cudaError_t cudaFuncSetAttribute(const void *f, enum cudaFuncAttribute a, int v) { (void) f; (void) a; (void) v; return cudaSuccess; }

/* peer memory is "not available": a Motion would take the staged path (none of these tests runs one) */
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { (void) h; (void) p; return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned int f) { (void) p; (void) h; (void) f; return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void *p) { (void) p; return cudaErrorNotSupported; }

/* a launch = nothing happens on the "device" */
cudaError_t
cudaLaunchKernel(const void *f, dim3 g, dim3 b, void **args, size_t shmem, cudaStream_t s)
{
	(void) f; (void) g; (void) b; (void) args; (void) shmem; (void) s;
	launches++;
	return cudaSuccess;
}

/* what nvcc's host stubs call around a <<< >>> launch and at load time */
static struct { dim3 g, b; size_t shmem; void *stream; } cfg;

unsigned
__cudaPushCallConfiguration(dim3 g, dim3 b, size_t shmem, void *stream)
{
	cfg.g = g; cfg.b = b; cfg.shmem = shmem; cfg.stream = stream;
	return 0;
}

cudaError_t
__cudaPopCallConfiguration(dim3 *g, dim3 *b, size_t *shmem, void *stream)
{
	*g = cfg.g; *b = cfg.b; *shmem = cfg.shmem; *(void **) stream = cfg.stream;
	return cudaSuccess;
}

static void *fatbin_handle;
void	  **__cudaRegisterFatBinary(void *fatCubin) { (void) fatCubin; return &fatbin_handle; }
void		__cudaRegisterFatBinaryEnd(void **h) { (void) h; }
void		__cudaUnregisterFatBinary(void **h) { (void) h; }
void		__cudaRegisterFunction(void **h, const char *hostFun, char *deviceFun, const char *deviceName, int thread_limit,
								   void *tid, void *bid, void *bDim, void *gDim, int *wSize)
{
	(void) h; (void) hostFun; (void) deviceFun; (void) deviceName; (void) thread_limit; (void) tid; (void) bid; (void) bDim; (void) gDim; (void) wSize;
}
void		__cudaRegisterVar(void **h, char *hostVar, char *deviceAddress, const char *deviceName, int ext, size_t size, int constant, int global)
{
	(void) h; (void) hostVar; (void) deviceAddress; (void) deviceName; (void) ext; (void) size; (void) constant; (void) global;
}
