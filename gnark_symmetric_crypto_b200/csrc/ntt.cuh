// Fr number-theoretic transforms and the Groth16 quotient H = (A.B - C)/Z_H for sm_100a — device kernels.
//
// Replaces: gnark-crypto v0.14.0 ecc/bn254/fr/fft/fft.go:42-407 ((*Domain).FFT / FFTInverse,
// difFFT, ditFFT, kerDIF*/kerDIT*) and gnark v0.11.0 backend/groth16/bn254/prove.go:359-384 (computeH), reached from
// libraries/prover/impl/provers.go:148,216.
//
// A transform of size n = 2^k is run as 1 + ceil((k-T)/8) passes (T = NTT_TILE_LG = 10); each pass stages a tile of up to
// 2^T elements (32 KiB) in shared memory and performs all radix-2 stages whose butterfly bit lies inside the tile, so a 2^15 vector
// crosses HBM/L2 twice per transform. DIF (natural -> bit-reversed) and DIT (bit-reversed -> natural) are paired exactly
// as gnark pairs them, which removes every bit-reversal permutation from compute_h; the 1/n and coset factors g^(+-i)
// are folded into the pass that touches the bit-reversed side. `batch` vectors are transformed per launch (grid.y).
#pragma once
#include "ntt_api.hpp"

namespace g16 {

#ifndef NTT_MINB
#define NTT_MINB 4   // resident CTAs per SM the pass kernel is compiled for (32 KiB tiles: 4 x 256 threads x 64 registers)
#endif
#ifndef NTT_RADIX4
#define NTT_RADIX4 1   // two stages per shared-memory round trip in registers (see the kernel). Measured on B200, 1024 proofs:
                       // compute_h 27.76 ms (radix 2) -> 27.37 ms (this, 4 CTAs / 64 registers with 60 bytes of spills) / 27.51 ms
                       // (3 CTAs / 80 registers): half the LDS / STS and barriers buy 1.4 % — the pass is bound by the multiplier
#endif
#ifndef NTT_ILP
#define NTT_ILP 1   // independent butterflies in flight per thread (see ntt_api.hpp for the measurements)
#endif

// tile-local index -> index in the vector
FD uint32_t ntt_global_index(const NttPass& p, uint32_t tile, uint32_t e) {
    if (p.b_lo == 0) return (tile << p.lg_tile) | e;
    // strided: e = (j << q) | l ; bits [q, b_lo) and bits >= b_lo+m come from the tile id
    uint32_t l = e & ((1u << p.q) - 1u), j = e >> p.q;
    uint32_t mid_bits = p.b_lo - p.q;
    uint32_t mid = tile & ((1u << mid_bits) - 1u), hi = tile >> mid_bits;
    return (hi << (p.b_lo + p.m)) | (j << p.b_lo) | (mid << p.q) | l;
}

// shared-memory layout: two planes of 16-byte halves, so consecutive elements are consecutive 16 B words (conflict-free
// LDS.128 for unit-stride access)
// (Measured on B200 and not adopted: an XOR swizzle e ^ 7 on elements with bit 3 set removes the 2-way bank conflicts of
// the stages on tile bits 0..2 — 21 % of the shared wavefronts are replays — but its index arithmetic costs more issue slots
// than the replays: compute_h 41.1 vs 40.3 ms per 1024 proofs. The kernel is bound by the multiplier pipe, not by LDS.)
FD void sm_store(uint4* sm, uint32_t tile_elems, uint32_t e, const Fr& v) {
    sm[e] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    sm[tile_elems + e] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
FD Fr sm_load(const uint4* sm, uint32_t tile_elems, uint32_t e) {
    uint4 a = sm[e], b = sm[tile_elems + e];
    Fr v;
    v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w;
    v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
    return v;
}

// One pass over `batch` vectors. tw = table of w^i, i < n/2 (w = the root used by this direction).
// scale (optional) multiplies element at vector index i by scale[i]: on load for DIT, on store for DIF.
__global__ void __launch_bounds__(NTT_THREADS, NTT_MINB)
ntt_pass_kernel(Fr* __restrict__ data, size_t vec_stride, NttPass p, const Fr* __restrict__ tw,
                const Fr* __restrict__ scale) {
#if defined(G16_EMU)
    uint4* sm = reinterpret_cast<uint4*>(cuemu::g_dyn_smem);
#else
    extern __shared__ uint4 sm[];
#endif
    const uint32_t tile_elems = 1u << p.lg_tile;
    const uint32_t tile = blockIdx.x;
    Fr* vec = data + (size_t)blockIdx.y * vec_stride;
    for (uint32_t e = threadIdx.x; e < tile_elems; e += blockDim.x) {
        uint32_t gi = ntt_global_index(p, tile, e);
        Fr v = vec[gi];
        if (scale && !p.dif) v = v * scale[gi];
        sm_store(sm, tile_elems, e, v);
    }
    __syncthreads();
    const uint32_t half = tile_elems >> 1;
    int s_first = 0;
#if NTT_RADIX4
    // Two stages per shared-memory round trip: a thread owns the four elements that differ in the two butterfly bits, runs both
    // stages in registers (4 products, 3 distinct twiddles) and stores once: half the LDS / STS traffic and half the barriers
    // of the radix-2 loop below, which keeps the odd last stage.
    for (; s_first + 1 < p.m; s_first += 2) {
        const int bitA = p.dif ? (p.b_lo + p.m - 1 - s_first) : (p.b_lo + s_first);        // stage done first
        const int bitB = p.dif ? bitA - 1 : bitA + 1;                                      // stage done second
        const int off = p.b_lo == 0 ? 0 : p.q;
        const int lbA = bitA - p.b_lo + off, lbB = bitB - p.b_lo + off;
        const int lo = lbA < lbB ? lbA : lbB;
        for (uint32_t u = threadIdx.x; u < (tile_elems >> 2); u += blockDim.x) {
            const uint32_t e00 = ((u >> lo) << (lo + 2)) | (u & ((1u << lo) - 1u));
            const uint32_t eA = 1u << lbA, eB = 1u << lbB;                                  // local index bits of the two stages
            Fr x00 = sm_load(sm, tile_elems, e00), xA = sm_load(sm, tile_elems, e00 | eA);
            Fr xB = sm_load(sm, tile_elems, e00 | eB), xAB = sm_load(sm, tile_elems, e00 | eA | eB);
            const uint32_t g00 = ntt_global_index(p, tile, e00), gB = ntt_global_index(p, tile, e00 | eB), gA = ntt_global_index(p, tile, e00 | eA);
            // stage A pairs (x00, xA) and (xB, xAB); stage B pairs (x00, xB) and (xA, xAB)
            if (p.dif) {
                Fr t0 = x00 + xA, t1 = x00 - xA, t2 = xB + xAB, t3 = xB - xAB;
                if (bitA != 0) {
                    t1 = t1 * tw[(g00 & ((1u << bitA) - 1u)) << (p.k - 1 - bitA)];
                    t3 = t3 * tw[(gB & ((1u << bitA) - 1u)) << (p.k - 1 - bitA)];
                }
                x00 = t0 + t2; xB = t0 - t2; xA = t1 + t3; xAB = t1 - t3;
                if (bitB != 0) {   // both pairs share the twiddle: their indices differ in bitA only, which lies above bitB
                    const Fr w = tw[(g00 & ((1u << bitB) - 1u)) << (p.k - 1 - bitB)];
                    xB = xB * w; xAB = xAB * w;
                }
            } else {
                if (bitA != 0) {   // shared twiddle of the first stage: the pairs differ in bitB only, above bitA
                    const Fr w = tw[(g00 & ((1u << bitA) - 1u)) << (p.k - 1 - bitA)];
                    xA = xA * w; xAB = xAB * w;
                }
                Fr t0 = x00 + xA, t1 = x00 - xA, t2 = xB + xAB, t3 = xB - xAB;   // t0: e00, t1: e00|A, t2: e00|B, t3: e00|A|B
                t2 = t2 * tw[(g00 & ((1u << bitB) - 1u)) << (p.k - 1 - bitB)];
                t3 = t3 * tw[(gA & ((1u << bitB) - 1u)) << (p.k - 1 - bitB)];
                x00 = t0 + t2; xB = t0 - t2; xA = t1 + t3; xAB = t1 - t3;
            }
            sm_store(sm, tile_elems, e00, x00);
            sm_store(sm, tile_elems, e00 | eA, xA);
            sm_store(sm, tile_elems, e00 | eB, xB);
            sm_store(sm, tile_elems, e00 | eA | eB, xAB);
        }
        __syncthreads();
    }
#endif
    for (int s = s_first; s < p.m; s++) {
        int bit = p.dif ? (p.b_lo + p.m - 1 - s) : (p.b_lo + s);   // butterfly bit in the vector index
        int lb = bit - p.b_lo + (p.b_lo == 0 ? 0 : p.q);           // same bit in the tile-local index
        // NTT_ILP butterflies per thread are loaded, computed and stored together: their Montgomery products are independent
        // carry chains that ptxas interleaves (one chain per warp leaves the multiplier waiting on its own carries)
        for (uint32_t u0 = threadIdx.x; u0 < half; u0 += blockDim.x * NTT_ILP) {
            Fr a[NTT_ILP], b[NTT_ILP], w[NTT_ILP];
            uint32_t e0[NTT_ILP];
#pragma unroll
            for (int j = 0; j < NTT_ILP; j++) {
                uint32_t u = u0 + (uint32_t)j * blockDim.x;
                if (u >= half) { e0[j] = 0xFFFFFFFFu; continue; }
                e0[j] = ((u >> lb) << (lb + 1)) | (u & ((1u << lb) - 1u));
                a[j] = sm_load(sm, tile_elems, e0[j]);
                b[j] = sm_load(sm, tile_elems, e0[j] | (1u << lb));
                if (bit != 0) {
                    uint32_t gi = ntt_global_index(p, tile, e0[j]);
                    w[j] = tw[(gi & ((1u << bit) - 1u)) << (p.k - 1 - bit)];
                }
            }
#pragma unroll
            for (int j = 0; j < NTT_ILP; j++) {
                if (e0[j] == 0xFFFFFFFFu) continue;
                Fr lo, hi;
                if (bit == 0) {   // span-1 stage: every twiddle is w^0 = 1 (uniform across the grid): no product
                    lo = a[j] + b[j];
                    hi = a[j] - b[j];
                } else if (p.dif) {
                    lo = a[j] + b[j];
                    hi = (a[j] - b[j]) * w[j];
                } else {
                    Fr t = b[j] * w[j];
                    lo = a[j] + t;
                    hi = a[j] - t;
                }
                sm_store(sm, tile_elems, e0[j], lo);
                sm_store(sm, tile_elems, e0[j] | (1u << lb), hi);
            }
        }
        __syncthreads();
    }
    for (uint32_t e = threadIdx.x; e < tile_elems; e += blockDim.x) {
        uint32_t gi = ntt_global_index(p, tile, e);
        Fr v = sm_load(sm, tile_elems, e);
        if (scale && p.dif) v = v * scale[gi];
        vec[gi] = v;
    }
}

// tw[i] = w^i for i < count (exponent has <= 32 bits)
__global__ void ntt_powers_kernel(Fr w, uint32_t count, Fr* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    Fr r = Fr::one();
    for (int b = 31; b >= 0; b--) {
        r = r.sqr();
        if ((i >> b) & 1u) r = r * w;
    }
    out[i] = r;
}
// out[j] = f * g^(brev_k(j))  for j < n
__global__ void ntt_coset_table_kernel(Fr g, Fr f, int k, Fr* __restrict__ out) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (1u << k)) return;
    uint32_t e = k ? (__brev(j) >> (32 - k)) : 0u;
    Fr r = Fr::one();
    for (int b = 31; b >= 0; b--) {
        r = r.sqr();
        if ((e >> b) & 1u) r = r * g;
    }
    out[j] = r * f;
}
__global__ void ntt_bitrev_kernel(const Fr* __restrict__ in, Fr* __restrict__ out, int k) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= (1u << k)) return;
    uint32_t r = k ? (__brev(j) >> (32 - k)) : 0u;
    out[r] = in[j];
}
// compute_h, pointwise steps (see compute_h_run): a[i] *= b[i] on the coset ; a[i] -= c[i] on the coefficients
__global__ void h_mul_kernel(Fr* __restrict__ a, const Fr* __restrict__ b, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    a[i] = a[i] * b[i];
}
__global__ void h_sub_kernel(Fr* __restrict__ a, const Fr* __restrict__ c, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    a[i] = a[i] - c[i];
}

// rows of unit vectors (evaluation-basis tables of the Z query, g16_ctx.cuh ctx_build_eval_tables): out must be zeroed;
// row r gets +-1 (Montgomery) at position j0 + r
__global__ void unit_rows_kernel(Fr* __restrict__ out, uint32_t n, uint32_t rows, uint32_t j0, int negate) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows || j0 + r >= n) return;
    Fr one = Fr::one();
    out[(size_t)r * n + j0 + r] = negate ? one.neg() : one;
}

// every constant is derived on the device: the product has no host-side field arithmetic
__global__ void ntt_domain_consts_kernel(Fr w, Fr g, uint32_t n, Fr* out /* [0]=w^-1 [1]=1/n [2]=g^-1 [3]=den [4]=den/n */) {
    if (threadIdx.x || blockIdx.x) return;
    out[0] = w.inv();
    Fr nn = Fr::zero();
    nn.l[0] = n;
    nn = nn.to_mont();
    out[1] = nn.inv();
    out[2] = g.inv();
    Fr gn = g;
    for (uint32_t m = 1; m < n; m <<= 1) gn = gn.sqr();
    out[3] = (gn - Fr::one()).inv();
    out[4] = out[3] * out[1];
}
// w = root28^(2^(28-k)), g = 5, both Montgomery
__global__ void ntt_root_kernel(const uint8_t* root_be, int k, Fr* out /* [0]=w [1]=g */) {
    if (threadIdx.x || blockIdx.x) return;
    Fr v;
    for (int i = 0; i < 8; i++) {
        const uint8_t* q = root_be + 28 - 4 * i;
        v.l[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
    v = v.to_mont();
    for (int i = 0; i < 28 - k; i++) v = v.sqr();
    out[0] = v;
    Fr one = Fr::one(), two = one + one;
    out[1] = two + two + one;
}
__global__ void fr_be_to_mont_kernel(const uint8_t* __restrict__ in, uint32_t n, Fr* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* b = in + (size_t)i * 32;
    Fr v;
    for (int k = 0; k < 8; k++) {
        const uint8_t* q = b + 28 - 4 * k;
        v.l[k] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
    for (int k = 0; k < 5; k++) v.reduce_once();
    out[i] = v.to_mont();
}
__global__ void fill_pattern_kernel(Fr* __restrict__ out, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    Fr v;
    uint32_t x = (uint32_t)i * 2654435761u + 12345u;
    for (int k = 0; k < 8; k++) { x ^= x << 13; x ^= x >> 17; x ^= x << 5; v.l[k] = x; }
    v.l[7] &= 0x0FFFFFFFu;   // < r
    out[i] = v;
}
__global__ void count_mismatch_kernel(const Fr* __restrict__ a, const Fr* __restrict__ b, size_t total,
                                      uint32_t* __restrict__ mismatches) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    if (a[i] != b[i]) atomicAdd(mismatches, 1u);
}

}  // namespace g16
