// TEST-ONLY host emulation of the small CUDA surface the kernels use, enabled with -DG16_EMU (built by tests/emu/Makefile
// into tests/emu/_build/libg16emu.so). It exists so that kernel LOGIC (index math, carry chains, sorting, reductions,
// host orchestration) can be exercised by the `-m "not gpu"` suite in a container without a GPU.
// It is NOT a fallback: the product loader (gnark_symmetric_crypto_b200/_lib.py) only ever loads the nvcc-built
// libg16b200.so and raises if it is missing; nothing in the package references the emulation library.
#pragma once
#include <atomic>
#include <barrier>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define __constant__ static

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return {x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return {x, y}; }

namespace cuemu {
extern thread_local dim3 t_threadIdx, t_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern std::barrier<>* g_barrier;
extern unsigned char* g_dyn_smem;
}  // namespace cuemu
#define threadIdx (cuemu::t_threadIdx)
#define blockIdx (cuemu::t_blockIdx)
#define blockDim (cuemu::g_blockDim)
#define gridDim (cuemu::g_gridDim)

static inline void __syncthreads() {
    if (cuemu::g_barrier) cuemu::g_barrier->arrive_and_wait();
}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __threadfence() {}

template <class T>
static inline T atomicAdd(T* p, T v) {
    return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
static inline uint32_t atomicOr(uint32_t* p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline uint32_t atomicAnd(uint32_t* p, uint32_t v) { return __atomic_fetch_and(p, v, __ATOMIC_RELAXED); }
static inline int atomicMax(int* p, int v) {
    int old = *p;
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline unsigned __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
template <class T> static inline T __ldg(const T* p) { return *p; }

// ---- launch: blocks run one after another; threads of a block run serially (sync == false) or as real threads
namespace cuemu {
template <class F>
void launch(dim3 grid, dim3 block, size_t smem, bool sync, F&& body) {
    g_blockDim = block;
    g_gridDim = grid;
    std::vector<unsigned char> dyn(smem + 16);
    g_dyn_smem = dyn.data();
    unsigned nthreads = block.x * block.y * block.z;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                if (!sync) {
                    g_barrier = nullptr;
                    t_blockIdx = dim3(bx, by, bz);
                    for (unsigned tz = 0; tz < block.z; tz++)
                        for (unsigned ty = 0; ty < block.y; ty++)
                            for (unsigned tx = 0; tx < block.x; tx++) {
                                t_threadIdx = dim3(tx, ty, tz);
                                body();
                            }
                } else {
                    std::barrier<> bar(nthreads);
                    g_barrier = &bar;
                    std::vector<std::thread> th;
                    for (unsigned t = 0; t < nthreads; t++)
                        th.emplace_back([&, t]() {
                            t_blockIdx = dim3(bx, by, bz);
                            t_threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                            body();
                            // threads that return early must not deadlock the others
                            g_barrier->arrive_and_drop();
                        });
                    for (auto& x : th) x.join();
                    g_barrier = nullptr;
                }
            }
}
}  // namespace cuemu

// ---- runtime shims
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef struct CuEmuEvent { double t; }* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? 0 : 2; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? 0 : 2; }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMallocHost((void**)p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaPeekAtLastError() { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new CuEmuEvent{0}; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = 0);
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)((b->t - a->t) * 1e3); return 0; }

#define G16_LAUNCH(kernel, grid, block, smem, stream, sync, ...) \
    cuemu::launch(dim3(grid), dim3(block), (smem), (sync), [&]() { kernel(__VA_ARGS__); })
