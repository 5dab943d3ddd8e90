// Fixed-base scalar multiplication by 4-bit window tables, shared by the G1 half (assemble.cuh, hot TU) and the G2 half
// (prover_kernels.cuh, cold TU) of the proof assembly. Replaces gnark-crypto v0.14.0 ecc/bn254
// BatchScalarMultiplicationG1(&pk.G1.Delta, {r, s, -rs}) and G2Jac.ScalarMultiplication(&pk.G2.Delta, s) as called from
// gnark v0.11.0 backend/groth16/bn254/prove.go:174-295.
#pragma once
#include "common.cuh"

namespace g16 {

static const int FB_WINDOWS = 64, FB_ENTRIES = 15;

// tab[i*15 + (j-1)] = j * 16^i * base  (affine). One thread per window.
template <class C>
__global__ void fixed_base_table_kernel(typename C::A base, typename C::A* __restrict__ tab) {
    typedef typename C::X X;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= FB_WINDOWS) return;
    X b = X::from_affine(base);
    for (int d = 0; d < 4 * i; d++) b = b.dbl();
    X acc = b;
    tab[i * FB_ENTRIES] = acc.to_affine();
    for (int j = 1; j < FB_ENTRIES; j++) {
        acc.add(b);
        tab[i * FB_ENTRIES + j] = acc.to_affine();
    }
}

FD uint32_t limb_of(const Fr& v, uint32_t i) {   // compile-time limb indices only (no local-memory copy of v)
    uint32_t r = 0;
#pragma unroll
    for (uint32_t k = 0; k < 8; k++) r = (i == k) ? v.l[k] : r;
    return r;
}

// k * base for a block of FB_WINDOWS threads: thread t fetches the table entry of nibble t of k (canonical limbs), the 64
// points meet in a shared-memory tree of depth 6. k * delta is therefore 6 dependent additions instead of a 64-step chain.
// sm: FB_WINDOWS entries. Every thread of the block must call; the result is valid in every thread.
template <class C>
__device__ __forceinline__ typename C::X fixed_base_mul_block(const typename C::A* __restrict__ tab, const Fr& k,
                                                              typename C::X* sm) {
    typedef typename C::X X;
    const uint32_t t = threadIdx.x;
    const uint32_t nib = (limb_of(k, t >> 3) >> (4 * (t & 7))) & 15u;
    sm[t] = nib ? X::from_affine(tab[t * FB_ENTRIES + (nib - 1)]) : X::inf();
    __syncthreads();
    for (uint32_t off = FB_WINDOWS / 2; off > 0; off >>= 1) {
        if (t < off) {
            X a = sm[t];
            a.add(sm[t + off]);
            sm[t] = a;
        }
        __syncthreads();
    }
    return sm[0];
}

FD Scalar256 scalar_of(const Fr& k) {
    Scalar256 s;
    for (int i = 0; i < 8; i++) s.w[i] = k.l[i];
    return s;
}

}  // namespace g16
