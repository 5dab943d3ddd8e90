// Proof assembly and serialisation, G1 half (the G2 element Bs is assembled by assemble_g2_kernel in prover_kernels.cuh,
// on the side stream that produced the G2 MSM). Replaces gnark v0.11.0 backend/groth16/bn254/prove.go:174-295 (tail of
// Prove) and marshal.go:32-59 (Proof.WriteTo), reached from libraries/prover/impl/provers.go:148-157 — SURVEY §8 a16, a18.
//
// 
//   Ar  = msmA + alpha + r*delta          Bs1 = msmB1 + beta + s*delta        Bs = msmB2 + beta2 + s*delta2
//   Krs = msmK + msmZ + s*Ar + r*Bs1 - (r*s)*delta
// The stage is pure latency (a handful of points per proof), so it is laid out for the single-request case:
//   phase 1  block = (proof, role in {r*delta, s*delta, -rs*delta}) x 64 threads: fixed-base products as a depth-6 tree
//   phase 2  warp roles {k1 Ar, k2 phi(Ar), k1' Bs1, k2' phi(Bs1), compress Ar}: the two variable-base products s*Ar, r*Bs1 are
//            the only long chains left; each is split with the GLV endomorphism into two 128-bit halves (4-bit windows: ~130
//            doublings + <= 33 additions per half), then Krs, its inversion and compression
// Measured on B200 for one request: the two variable-base products are a chain of ~3.3 k dependent Montgomery products at
// ~0.6 us each (2.0 ms of the 2.2 ms stage); inlining the product (this TU is "hot") or not makes no difference.
#pragma once
#include "fixed_base.cuh"
#include "glv_consts.hpp"
#include "team.cuh"
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

// out = low NO limbs of a (NA limbs) * b (NB limbs). Fully unrolled: every array stays in registers.
template <int NA, int NB, int NO>
FD void mul_limbs(const uint32_t* a, const uint32_t* b, uint32_t* out) {
#pragma unroll
    for (int i = 0; i < NO; i++) out[i] = 0;
#pragma unroll
    for (int i = 0; i < NA; i++) {
        uint64_t carry = 0;
#pragma unroll
        for (int j = 0; j < NB; j++) {
            if (i + j < NO) {
                uint64_t t = (uint64_t)a[i] * b[j] + out[i + j] + carry;
                out[i + j] = (uint32_t)t;
                carry = t >> 32;
            }
        }
        if (i + NB < NO) out[i + NB] = (uint32_t)carry;
    }
}
// k = k1 + k2 lambda (mod r) with |k1|, |k2| < 2^132 (glv_consts.hpp). ok = false if a half does not fit (never observed; the
// caller then falls back to the plain 254-bit product).
struct GlvSplit {
    uint32_t k1[5], k2[5];
    bool neg1, neg2, ok;
};
FD bool glv_abs(uint32_t v[8], bool& neg) {   // two's complement -> magnitude ; true if it fits in 132 bits
    neg = (v[7] >> 31) != 0;
    if (neg) {
        uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        sub8(v, z, v);
    }
    return (v[5] | v[6] | v[7]) == 0 && v[4] < 16u;
}
FD GlvSplit glv_split(const Fr& k) {   // k: canonical limbs
    uint32_t g1[3], g2[5], a1[2], a2[4], nb1[4], b2[2];
#pragma unroll
    for (int i = 0; i < 3; i++) g1[i] = GLV_G1[i];
#pragma unroll
    for (int i = 0; i < 5; i++) g2[i] = GLV_G2[i];
#pragma unroll
    for (int i = 0; i < 2; i++) { a1[i] = GLV_A1[i]; b2[i] = GLV_B2[i]; }
#pragma unroll
    for (int i = 0; i < 4; i++) { a2[i] = GLV_A2[i]; nb1[i] = GLV_NB1[i]; }
    uint32_t t[11], u[13];
    mul_limbs<8, 3, 11>(k.l, g1, t);    // c1 = t[8..10]
    mul_limbs<8, 5, 13>(k.l, g2, u);    // c2 = u[8..12]
    uint32_t p1[8], p2[8], q1[8], q2[8], v1[8], v2[8];
    mul_limbs<3, 2, 8>(t + 8, a1, p1);
    mul_limbs<5, 4, 8>(u + 8, a2, p2);
    mul_limbs<3, 4, 8>(t + 8, nb1, q1);
    mul_limbs<5, 2, 8>(u + 8, b2, q2);
    sub8(v1, k.l, p1);
    sub8(v1, v1, p2);       // k1 = k - c1 a1 - c2 a2   (mod 2^256, the true value is small)
    sub8(v2, q1, q2);       // k2 = c1 |b1| - c2 b2
    GlvSplit r;
    bool ok1 = glv_abs(v1, r.neg1), ok2 = glv_abs(v2, r.neg2);
    r.ok = ok1 && ok2;
#pragma unroll
    for (int i = 0; i < 5; i++) { r.k1[i] = v1[i]; r.k2[i] = v2[i]; }
    return r;
}

// k * P with 4-bit windows, k given as NL limbs of which the low `nibbles` 4-bit digits are used; tab = 15 entries of scratch
// owned by this thread (global memory, L1/L2 resident), entry j at tab[j * ts]
template <int NL>
__device__ __forceinline__ G1XYZZ window_mul(const G1XYZZ& P, const uint32_t* k, int nibbles, G1XYZZ* __restrict__ tab, size_t ts) {
    G1XYZZ acc = P;
    tab[0] = P;
    for (int j = 1; j < 15; j++) {
        if (j & 1) acc = tab[(size_t)(j >> 1) * ts].dbl();   // (j+1) P = 2 * ((j+1)/2) P
        else acc.add(P);                                        // j+1 odd: previous (even multiple) + P
        tab[(size_t)j * ts] = acc;
    }
    acc = G1XYZZ::inf();
#pragma unroll
    for (int wi = NL - 1; wi >= 0; wi--) {
        const uint32_t word = k[wi];
#pragma unroll 1
        for (int j = 7; j >= 0; j--) {
            if (wi * 8 + j >= nibbles) continue;
#pragma unroll 1   // one copy of the doubling: unrolled x4 the loop body overflows the instruction cache (2.7 vs 2.2 ms)
            for (int d = 0; d < 4; d++) acc = acc.dbl();
            const uint32_t nib = (word >> (4 * j)) & 15u;
            if (nib) acc.add(tab[(size_t)(nib - 1) * ts]);
        }
    }
    return acc;
}

// grid ceil(n / 32), block (32, 5). s*Ar and r*Bs1 are split with the GLV endomorphism phi(x, y) = (beta x, y) = lambda (x, y):
// k P = k1 P + k2 phi(P) with 128-bit halves, so the longest chain is ~130 doublings instead of 254.
// warp 0 = k1 Ar (then Krs) ; 1 = k2 phi(Ar) ; 2 = k1' Bs1 ; 3 = k2' phi(Bs1) ; 4 = Ar compressed + proof trailer
__global__ void __launch_bounds__(160)
assemble_phase2_kernel(uint32_t n, int with_commitment, const G1XYZZ* __restrict__ mK, const G1XYZZ* __restrict__ mZ,
                       const G1XYZZ* __restrict__ Ar, const G1XYZZ* __restrict__ Bs1, const G1XYZZ* __restrict__ nrsd,
                       const Fr* __restrict__ rs, G1XYZZ* __restrict__ win_tab, uint8_t* __restrict__ out, size_t out_stride) {
    __shared__ G1XYZZ part[3][32];
    const uint32_t i = blockIdx.x * 32 + threadIdx.x, role = threadIdx.y;
    const bool live = i < n;
    G1XYZZ acc = G1XYZZ::inf();
    if (live && role < 4) {
        // window tables: entry j of (proof i, role) at win_tab[(j * 4 + role) * n + i]  (coalesced across the warp)
        G1XYZZ* tab = win_tab + (size_t)role * n + i;
        const size_t ts = (size_t)4 * n;
        const bool second = (role & 1u) != 0;                 // the phi half
        G1XYZZ P = role < 2 ? Ar[i] : Bs1[i];
        const Fr k = role < 2 ? rs[2 * i + 1] : rs[2 * i];     // s for Ar, r for Bs1
        const GlvSplit sp = glv_split(k);
        G1XYZZ v;
        if (sp.ok) {
            if (second) {
                Fp beta;
#pragma unroll
                for (int q = 0; q < 8; q++) beta.l[q] = GLV_BETA[q];
                P.X = P.X * beta.to_mont();
            }
            if (second ? sp.neg2 : sp.neg1) P = P.neg();
            v = window_mul<5>(P, second ? sp.k2 : sp.k1, GLV_NIBBLES, tab, ts);
        } else {
            v = second ? G1XYZZ::inf() : window_mul<8>(P, k.l, 64, tab, ts);
        }
        if (role == 0) acc = v;
        else part[role - 1][threadIdx.x] = v;
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
    acc.add(part[0][threadIdx.x]);
    acc.add(part[1][threadIdx.x]);
    acc.add(part[2][threadIdx.x]);
    acc.add(nrsd[i]);
    acc.add(mK[i]);
    acc.add(mZ[i]);
    g1_compress(acc.to_affine(), out + (size_t)i * out_stride + 96);
}

// ---- the same step with the four half-products run by teams of four warps (team.cuh). A thread alone needs 3.95 us per
// doubling and 6.3 us per addition (its partition's multiplier serves one live lane at the price of 32), and phase 2 is ~130
// doublings + ~45 additions deep whatever the batch: 1.18 ms for 1 proof or for 1024, on 32 of the 148 SMs. Here every
// (32 proofs, role) is one 128-thread block = one team: grid (ceil(n / 32), 4), the independent products of each formula go one
// to each warp (2.5 / 3.9 us per doubling / addition), and a second short kernel adds the four halves to the rest of Krs.
// role 0 = k1 Ar ; 1 = k2 phi(Ar) ; 2 = k1' Bs1 ; 3 = k2' phi(Bs1). Lane l = proof 32 b + l; the 15-entry window table of a
// (proof, role) sits in win_tab as before (written by warp 0, read by all four after a barrier); the result goes to entry 15.
__global__ void __launch_bounds__(128)
assemble_mul_team_kernel(uint32_t n, const G1XYZZ* __restrict__ Ar, const G1XYZZ* __restrict__ Bs1, const Fr* __restrict__ rs,
                         G1XYZZ* win_tab) {
    __shared__ Fp sm[TEAM4_SM_ELEMS];
    __shared__ uint32_t flag;
    Team4 T{sm, &flag, (int)(threadIdx.x >> 5), (int)(threadIdx.x & 31), 0};
    if (threadIdx.x == 0) flag = 0u;
    __syncthreads();
    const uint32_t role = blockIdx.y;
    const uint32_t i = blockIdx.x * 32 + T.lane;
    const bool live = i < n;
    const bool second = (role & 1u) != 0;
    const size_t ts = (size_t)4 * n;
    G1XYZZ* tab = win_tab + (size_t)role * n + (live ? i : 0);
    G1XYZZ P = G1XYZZ::inf();
    uint32_t k[5] = {0, 0, 0, 0, 0};
    bool ok = true;
    if (live) {
        P = role < 2 ? Ar[i] : Bs1[i];
        const Fr kk = role < 2 ? rs[2 * i + 1] : rs[2 * i];     // s for Ar, r for Bs1
        const GlvSplit sp = glv_split(kk);
        ok = sp.ok;
        if (ok) {
            if (second) {
                Fp beta;
#pragma unroll
                for (int q = 0; q < 8; q++) beta.l[q] = GLV_BETA[q];
                P.X = P.X * beta.to_mont();
            }
            if (second ? sp.neg2 : sp.neg1) P = P.neg();
#pragma unroll
            for (int q = 0; q < 5; q++) k[q] = second ? sp.k2[q] : sp.k1[q];
        } else {
            P = G1XYZZ::inf();   // this lane takes the plain 254-bit product after the loop (never observed)
        }
    }
    // window table 1 P .. 15 P
    G1XYZZ acc = P;
    if (live && T.w == 0) tab[0] = acc;
    acc = team_dbl(T, P);
    if (live && T.w == 0) tab[ts] = acc;
    for (int j = 2; j < 15; j++) {
        acc = team_add(T, acc, P, false);
        if (live && T.w == 0) tab[(size_t)j * ts] = acc;
    }
    __syncthreads();
    acc = G1XYZZ::inf();
    // nibble 32 (bits 128..131) first; the five words shift left by one nibble per round (constant register indices only)
    for (int nb = GLV_NIBBLES - 1; nb >= 0; nb--) {
        if (nb != GLV_NIBBLES - 1)
            for (int d = 0; d < 4; d++) acc = team_dbl(T, acc);
        const uint32_t nib = k[4] & 15u;
        k[4] = (k[4] << 4) | (k[3] >> 28); k[3] = (k[3] << 4) | (k[2] >> 28); k[2] = (k[2] << 4) | (k[1] >> 28);
        k[1] = (k[1] << 4) | (k[0] >> 28); k[0] <<= 4;
        const G1XYZZ e = (live && nib) ? tab[(size_t)(nib - 1) * ts] : G1XYZZ::inf();
        acc = team_add(T, acc, e, nib == 0);
    }
    if (!live || T.w != 0) return;
    if (!ok) {
        const G1XYZZ Q = role < 2 ? Ar[i] : Bs1[i];
        const Fr kk = role < 2 ? rs[2 * i + 1] : rs[2 * i];
        acc = second ? G1XYZZ::inf() : window_mul<8>(Q, kk.l, 64, tab, ts);
    }
    tab[(size_t)15 * ts] = acc;
}
// block (32, 2): y = 0 sums Krs = the four halves - r s delta + K + Z ; y = 1 writes Ar and the proof trailer
__global__ void __launch_bounds__(64)
assemble_finish_kernel(uint32_t n, int with_commitment, const G1XYZZ* __restrict__ mK, const G1XYZZ* __restrict__ mZ,
                       const G1XYZZ* __restrict__ Ar, const G1XYZZ* __restrict__ nrsd, const G1XYZZ* __restrict__ win_tab,
                       uint8_t* __restrict__ out, size_t out_stride) {
    const uint32_t i = blockIdx.x * 32 + threadIdx.x;
    if (i >= n) return;
    if (threadIdx.y == 0) {
        const G1XYZZ* res = win_tab + (size_t)15 * 4 * n + i;
        G1XYZZ acc = res[0];
        acc.add(res[n]);
        acc.add(res[(size_t)2 * n]);
        acc.add(res[(size_t)3 * n]);
        acc.add(nrsd[i]);
        acc.add(mK[i]);
        acc.add(mZ[i]);
        g1_compress(acc.to_affine(), out + (size_t)i * out_stride + 96);
    } else {
        uint8_t* o = out + (size_t)i * out_stride;
        g1_compress(Ar[i].to_affine(), o);
        if (!with_commitment) {   // u32 0 commitments | infinity PoK (Appendix C)
            o[128] = 0; o[129] = 0; o[130] = 0; o[131] = 0;
            o[132] = 0x40;
            for (int k = 133; k < 164; k++) o[k] = 0;
        }
    }
}

}  // namespace g16
