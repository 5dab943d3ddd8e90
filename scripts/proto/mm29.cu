// Prototype: carry-free radix-2^29 Montgomery product for BN254 Fp (9 limbs, 64-bit column accumulators, R' = 2^261)
// versus the 8x32 carry-chain product of field.cuh. Build: nvcc -arch=sm_100a -O3 -I../../gnark_symmetric_crypto_b200/csrc
#include <cstdio>
#include <cstdint>
#include <vector>
#include "field.cuh"
using namespace g16;

#ifndef VARIANT
#define VARIANT 0
#endif
#define MASK29 0x1fffffffu
// p in radix 2^29
__device__ __constant__ uint32_t P29[9];
__device__ __constant__ uint32_t PINV29;   // -p^-1 mod 2^29

__device__ __forceinline__ void madw(uint64_t& acc, uint32_t a, uint32_t b) {
#if VARIANT == 1
    uint32_t lo = (uint32_t)acc, hi = (uint32_t)(acc >> 32);
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
    acc = ((uint64_t)hi << 32) | lo;
#else
    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
#endif
}
__device__ __forceinline__ uint64_t mulw(uint32_t a, uint32_t b) {
    uint64_t r;
    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
}

struct F29 { uint32_t l[9]; };

// c = a*b*2^-261 mod p (result < 2^255 + p, limbs < 2^29 except possibly nothing: fully carried)
template <bool CONSTP>
__device__ __forceinline__ F29 mul29(const F29& a, const F29& b, const uint32_t* p29, uint32_t pinv) {
    uint64_t t[18];
#pragma unroll
    for (int j = 0; j < 9; j++) t[j] = mulw(a.l[0], b.l[j]);
#pragma unroll
    for (int j = 9; j < 18; j++) t[j] = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        if (i > 0) {
#pragma unroll
            for (int j = 0; j < 9; j++) madw(t[i + j], a.l[i], b.l[j]);
        }
        uint32_t m = ((uint32_t)t[i] * pinv) & MASK29;
#pragma unroll
        for (int j = 0; j < 9; j++) madw(t[i + j], m, p29[j]);
        t[i + 1] += t[i] >> 29;
    }
    F29 r;
#pragma unroll
    for (int k = 9; k < 17; k++) {
        t[k + 1] += t[k] >> 29;
        r.l[k - 9] = (uint32_t)t[k] & MASK29;
    }
    r.l[8] = (uint32_t)t[17];
    return r;
}

__device__ __forceinline__ F29 to29(const Fp& x) {
    F29 r;
    uint64_t w[4] = {((uint64_t)x.l[1] << 32) | x.l[0], ((uint64_t)x.l[3] << 32) | x.l[2], ((uint64_t)x.l[5] << 32) | x.l[4],
                     ((uint64_t)x.l[7] << 32) | x.l[6]};
#pragma unroll
    for (int k = 0; k < 9; k++) {
        int bit = 29 * k, wi = bit >> 6, sh = bit & 63;
        uint64_t v = w[wi] >> sh;
        if (sh > 35 && wi < 3) v |= w[wi + 1] << (64 - sh);
        r.l[k] = (uint32_t)v & MASK29;
    }
    return r;
}
__device__ __forceinline__ Fp from29(const F29& x) {
    uint64_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 9; k++) {
        int bit = 29 * k, wi = bit >> 6, sh = bit & 63;
        w[wi] |= (uint64_t)x.l[k] << sh;
        if (sh > 35 && wi < 3) w[wi + 1] |= (uint64_t)x.l[k] >> (64 - sh);
    }
    Fp r;
#pragma unroll
    for (int k = 0; k < 4; k++) { r.l[2 * k] = (uint32_t)w[k]; r.l[2 * k + 1] = (uint32_t)(w[k] >> 32); }
    return r;
}

__global__ void __launch_bounds__(256) mm29_kernel(Fp* out, int iters, int check) {
    Fp a0 = Fp::one(), b0 = Fp::r2(), c0 = Fp::one(), d0 = Fp::r2();
    a0.l[0] += threadIdx.x; b0.l[0] += blockIdx.x; c0.l[1] += threadIdx.x; d0.l[1] += blockIdx.x;
    F29 a = to29(a0), b = to29(b0), c = to29(c0), d = to29(d0);
    uint32_t p29[9];
#pragma unroll
    for (int j = 0; j < 9; j++) p29[j] = P29[j];
    uint32_t pinv = PINV29;
    for (int i = 0; i < iters; i++) {
        a = mul29<false>(a, b, p29, pinv);
        c = mul29<false>(c, d, p29, pinv);
        b = mul29<false>(b, a, p29, pinv);
        d = mul29<false>(d, c, p29, pinv);
    }
    if (check) {
        size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        out[4 * g] = from29(a); out[4 * g + 1] = from29(b); out[4 * g + 2] = from29(c); out[4 * g + 3] = from29(d);
    } else if (a.l[0] == 0x12345u && b.l[3] == 77u && c.l[1] == d.l[2]) out[0] = from29(a);
}
__global__ void __launch_bounds__(256) mm32_kernel(Fp* out, int iters, int check) {
    Fp a = Fp::one(), b = Fp::r2(), c = Fp::one(), d = Fp::r2();
    a.l[0] += threadIdx.x; b.l[0] += blockIdx.x; c.l[1] += threadIdx.x; d.l[1] += blockIdx.x;
    for (int i = 0; i < iters; i++) {
        a = a * b; c = c * d; b = b * a; d = d * c;
    }
    if (check) {
        size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        out[4 * g] = a; out[4 * g + 1] = b; out[4 * g + 2] = c; out[4 * g + 3] = d;
    } else if (a.l[0] == 0x12345u && b.l[3] == 77u && c.l[1] == d.l[2]) out[0] = a;
}

int main() {
    // constants
    const uint32_t p32[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    uint32_t p29[9];
    unsigned __int128 dummy = 0; (void)dummy;
    for (int k = 0; k < 9; k++) {
        int bit = 29 * k;
        uint64_t v = 0;
        for (int b = 0; b < 29; b++) {
            int pos = bit + b;
            if (pos < 256 && ((p32[pos >> 5] >> (pos & 31)) & 1)) v |= 1ull << b;
        }
        p29[k] = (uint32_t)v;
    }
    // pinv = -p^-1 mod 2^29 by Newton
    uint32_t inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - p29[0] * inv;
    uint32_t pinv = (0u - inv) & MASK29;
    cudaMemcpyToSymbol(P29, p29, sizeof p29);
    cudaMemcpyToSymbol(PINV29, &pinv, 4);
    // correctness: 1 block x 256 threads, 3 iterations; print thread 5 results of both kernels as hex for the host script
    Fp* d; cudaMalloc(&d, 148 * 8 * 256 * 4 * sizeof(Fp));
    std::vector<Fp> h29(1024), h32(1024);
    mm29_kernel<<<1, 256>>>(d, 1, 1); cudaMemcpy(h29.data(), d, 1024 * sizeof(Fp), cudaMemcpyDeviceToHost);
    mm32_kernel<<<1, 256>>>(d, 1, 1); cudaMemcpy(h32.data(), d, 1024 * sizeof(Fp), cudaMemcpyDeviceToHost);
    for (int t = 0; t < 3; t++) {
        for (int k = 0; k < 4; k++) {
            printf("R29 %d %d ", t, k); for (int i = 7; i >= 0; i--) printf("%08x", h29[4 * (t * 37 + 1) + k].l[i]); printf("\n");
            printf("R32 %d %d ", t, k); for (int i = 7; i >= 0; i--) printf("%08x", h32[4 * (t * 37 + 1) + k].l[i]); printf("\n");
        }
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, threads = 256, iters = 400;
    for (int which = 0; which < 2; which++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            if (which == 0) mm32_kernel<<<blocks, threads>>>(d, iters, 0); else mm29_kernel<<<blocks, threads>>>(d, iters, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
        }
        printf("%s: %.3f ms  %.2f Gmul/s\n", which ? "radix29" : "radix32", best, (double)blocks * threads * iters * 4 / best / 1e6);
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
