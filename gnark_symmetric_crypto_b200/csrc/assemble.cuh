// Proof assembly and serialisation, G1 half (the G2 element Bs is assembled by assemble_g2_kernel in prover_kernels.cuh,
// on the side stream that produced the G2 MSM). Replaces gnark v0.11.0 backend/groth16/bn254/prove.go:174-295 (tail of
// Prove) and marshal.go:32-59 (Proof.WriteTo), reached from libraries/prover/impl/provers.go:148-157 — SURVEY §8 a16, a18.
//
// 
//   Ar  = msmA + alpha + r*delta          Bs1 = msmB1 + beta + s*delta        Bs = msmB2 + beta2 + s*delta2
//   Krs = msmK + msmZ + s*Ar + r*Bs1 - (r*s)*delta
// The stage is pure latency (a handful of points per proof), so it is laid out for the single-request case:
//   phase 1  block = (proof, role in {r*delta, s*delta, -rs*delta}) x 64 threads: fixed-base products as a depth-6 tree
//   phase 2  warp roles {s*Ar, r*Bs1, compress Ar}: the two variable-base products are the only long chains left
//            (4-bit windows: 252 doublings + <= 63 additions each), then Krs, its inversion and compression
// Measured on B200 for one request: the two variable-base products are a chain of ~3.3 k dependent Montgomery products at
// ~0.6 us each (2.0 ms of the 2.2 ms stage); inlining the product (this TU is "hot") or not makes no difference.
#pragma once
#include "fixed_base.cuh"
#include "prover_api.hpp"
#include "serialize.cuh"

namespace g16 {

// rs: canonical limbs, r at [2i], s at [2i+1].  grid (n, 3), block 64
__global__ void __launch_bounds__(FB_WINDOWS)
assemble_phase1_kernel(AssemblyKeys keys, uint32_t n, const G1XYZZ* __restrict__ mA, const G1XYZZ* __restrict__ mB1,
                       const Fr* __restrict__ rs, G1XYZZ* __restrict__ Ar_out, G1XYZZ* __restrict__ Bs1_out,
                       G1XYZZ* __restrict__ nrsd_out) {
    __shared__ G1XYZZ sm[FB_WINDOWS];
    const uint32_t i = blockIdx.x, role = blockIdx.y;
    Fr k;
    if (role == 0) k = rs[2 * i];
    else if (role == 1) k = rs[2 * i + 1];
    else k = (rs[2 * i].to_mont() * rs[2 * i + 1].to_mont()).from_mont();
    G1XYZZ v = fixed_base_mul_block<G1>(keys.delta_tab, k, sm);
    if (threadIdx.x) return;
    if (role == 0) {
        v.add(mA[i]);
        v.madd(keys.alpha, false);
        Ar_out[i] = v;
    } else if (role == 1) {
        v.add(mB1[i]);
        v.madd(keys.beta, false);
        Bs1_out[i] = v;
    } else {
        nrsd_out[i] = v.neg();
    }
}

// k * P with 4-bit windows; tab = 15 entries of scratch owned by this thread (global memory, L1/L2 resident)
__device__ __forceinline__ G1XYZZ window_mul(const G1XYZZ& P, const Fr& k, G1XYZZ* __restrict__ tab, size_t ts) {
    G1XYZZ acc = P;
    tab[0] = P;
    for (int j = 1; j < 15; j++) {
        if (j & 1) acc = tab[(size_t)(j >> 1) * ts].dbl();   // (j+1) P = 2 * ((j+1)/2) P
        else acc.add(P);                                        // j+1 odd: previous (even multiple) + P
        tab[(size_t)j * ts] = acc;
    }
    acc = G1XYZZ::inf();
#pragma unroll
    for (int wi = 7; wi >= 0; wi--) {
        const uint32_t word = k.l[wi];
#pragma unroll 1
        for (int j = 7; j >= 0; j--) {
#pragma unroll 1   // one copy of the doubling: unrolled x4 the loop body overflows the instruction cache (2.7 vs 2.2 ms)
            for (int d = 0; d < 4; d++) acc = acc.dbl();
            const uint32_t nib = (word >> (4 * j)) & 15u;
            if (nib) acc.add(tab[(size_t)(nib - 1) * ts]);
        }
    }
    return acc;
}

// grid ceil(n / 32), block (32, 3): warp 0 = s*Ar then Krs ; warp 1 = r*Bs1 ; warp 2 = Ar compressed + proof trailer
__global__ void __launch_bounds__(96)
assemble_phase2_kernel(uint32_t n, int with_commitment, const G1XYZZ* __restrict__ mK, const G1XYZZ* __restrict__ mZ,
                       const G1XYZZ* __restrict__ Ar, const G1XYZZ* __restrict__ Bs1, const G1XYZZ* __restrict__ nrsd,
                       const Fr* __restrict__ rs, G1XYZZ* __restrict__ win_tab, uint8_t* __restrict__ out, size_t out_stride) {
    __shared__ G1XYZZ rBs1[32];
    const uint32_t i = blockIdx.x * 32 + threadIdx.x, role = threadIdx.y;
    const bool live = i < n;
    G1XYZZ acc = G1XYZZ::inf();
    if (live && role < 2) {
        // window tables: entry j of (proof i, role) at win_tab[(j * 2 + role) * n + i]  (coalesced across the warp)
        G1XYZZ* tab = win_tab + (size_t)role * n + i;
        if (role == 0) acc = window_mul(Ar[i], rs[2 * i + 1], tab, (size_t)2 * n);
        else rBs1[threadIdx.x] = window_mul(Bs1[i], rs[2 * i], tab, (size_t)2 * n);
    } else if (live) {
        uint8_t* o = out + (size_t)i * out_stride;
        g1_compress(Ar[i].to_affine(), o);
        if (!with_commitment) {   // u32 0 commitments | infinity PoK (Appendix C)
            o[128] = 0; o[129] = 0; o[130] = 0; o[131] = 0;
            o[132] = 0x40;
            for (int k = 133; k < 164; k++) o[k] = 0;
        }
    }
    __syncthreads();
    if (!live || role != 0) return;
    acc.add(rBs1[threadIdx.x]);
    acc.add(nrsd[i]);
    acc.add(mK[i]);
    acc.add(mZ[i]);
    g1_compress(acc.to_affine(), out + (size_t)i * out_stride + 96);
}

}  // namespace g16
