// Shared host-side plumbing for the g16b200 library: error handling that never aborts across the C-ABI,
// launch macro, device buffers.
#pragma once
#include "field.cuh"
#include "ec.cuh"
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#if !defined(G16_EMU)
#define G16_LAUNCH(kernel, grid, block, smem, stream, sync, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

namespace g16 {

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define G16_CUDA(expr)                                                                                          \
    do {                                                                                                        \
        cudaError_t _e = (expr);                                                                                \
        if (_e != cudaSuccess)                                                                                  \
            throw g16::CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" +   \
                                 std::to_string(__LINE__) + ")");                                               \
    } while (0)

#define G16_CHECK_LAUNCH() G16_CUDA(cudaGetLastError())

// owning device buffer
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count) {
        release();
        if (count) G16_CUDA(cudaMalloc((void**)&p, count * sizeof(T)));   // on failure the buffer stays empty (n = 0)
        n = count;
    }
    void ensure(size_t count) { if (count > n) alloc(count); }
    void upload(const T* h, size_t count, cudaStream_t s = 0) {
        ensure(count);
        if (count) G16_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void download(T* h, size_t count, cudaStream_t s = 0) const {
        if (count) G16_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    }
    void zero(cudaStream_t s = 0) { if (n) G16_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    size_t bytes() const { return n * sizeof(T); }
};

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// streaming multiprocessors of the current device (148 on B200): grid sizing heuristics scale with it instead of naming it
static inline size_t sm_count() {
#if defined(G16_EMU)
    return 148;
#else
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return (size_t)cached[dev];
#endif
}

// CUDA-event stage timer: mark(stage) opens an interval attributed to `stage`; finish() (after the stream has been
// synchronised) adds every interval to its stage. Events are pooled and reused.
enum Stage { ST_SOLVE = 0, ST_H = 1, ST_MSM_SORT = 2, ST_MSM_ACC = 3, ST_MSM_REDUCE = 4, ST_ASSEMBLE = 5, ST_COUNT = 6 };
struct StageTimer {
    std::vector<cudaEvent_t> evs;
    std::vector<int> stage;
    size_t used = 0;
    bool enabled = true;
    ~StageTimer() { for (auto e : evs) cudaEventDestroy(e); }
    void reset() { used = 0; stage.clear(); }
    void mark(int st, cudaStream_t s) {
        if (!enabled) return;
        if (used == evs.size()) { cudaEvent_t e; G16_CUDA(cudaEventCreate(&e)); evs.push_back(e); }
        G16_CUDA(cudaEventRecord(evs[used], s));
        used++;
        stage.push_back(st);
    }
    // out[ST_COUNT]: per-stage sums in ms; returns the total. The last mark only closes the previous interval.
    float finish(float* out) {
        for (int i = 0; i < ST_COUNT; i++) out[i] = 0.f;
        float total = 0.f;
        for (size_t i = 0; i + 1 < used; i++) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, evs[i], evs[i + 1]);
            if (stage[i] >= 0 && stage[i] < ST_COUNT) out[stage[i]] += ms;
            total += ms;
        }
        return total;
    }
};

}  // namespace g16
