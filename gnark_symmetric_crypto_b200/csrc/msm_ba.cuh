// Batch-affine bucket accumulation for the G1 MSM: the first K levels of every bucket's sum are pairwise AFFINE additions that
// share one field inversion per block batch (Montgomery's trick), 6 products per addition instead of the 10 of the XYZZ mixed
// addition.
//
// Replaces: gnark-crypto v0.14.0 ecc/bn254/multiexp_affine.go:35-176 (processChunkG1BatchAffine: bucket
// additions queued and executed with one shared inversion) and g1.go:1133-1152 (batchAddG1Affine) — the algorithm gnark's own
// CPU MSM uses — reached from groth16 Prove (provers.go:148,216).
//
// Layout. The counting sort pads every non-empty bucket's run of entries to a multiple of 2^K slots (null references fill
// the tail), so runs start at multiples of 2^K and K rounds of "add slot 2q and slot 2q+1 into slot q" never pair entries of
// different buckets and need no keys. Level 0 reads point references into the base table, level l > 0 the points the level
// before wrote. After K levels one point per group of 2^K slots is left; those (1/2^K of the entries) go through the
// existing sorted-run XYZZ accumulation with one (bucket, group) key per group. The group sums are the same group elements
// whatever the association order, so bucket sums, MSM results and proof bytes do not change.
//
// One level = three kernels, none of which synchronises threads:
//   den   every thread walks its M pairs, forms the denominators d (x2 - x1, or 2 y1 when the two points are equal, or 1 when
//         the pair needs no division: a null / infinite operand, P + (-P)), stores the running products c_j = d_0 ... d_j
//         (32 bytes per pair) and its total c_(M-1)
//   inv   Montgomery's trick over the thread totals, 64 totals per thread: every lane of the machine runs one binary-Euclid
//         inversion at the same time (one inversion per 64 M additions)
//   add   every thread walks its pairs backwards: 1/d_j = c_(j-1) R, R <- R d_j, then lambda, x3, y3
// Products per pair: 1 (prefix) + 2 (back-substitution) + 3 (lambda, lambda^2, y3) + 3/M (totals) = 6.2 at M = 16.
// (Measured on B200 and not adopted: the same work as ONE persistent kernel with a block-wide product tree in shared memory
// and one inversion per block batch — the single-lane inversion costs a fifth of the kernel's issue slots and its latency
// keeps a block's other warps waiting: 38.2 ms for the three levels of a 512-proof Z query, 1.50 IPC.)
#pragma once
#include "msm_types.hpp"

namespace g16 {

static const uint32_t BA_NULL = 0xFFFFFFFFu;   // reference of a padding slot
// G16_BA_PREFETCH=1: the addition kernel requests the references and the running product of its next pair one iteration ahead.
// MEASURED ON B200 AND NOT THE DEFAULT: 92 instead of 88 registers, step 130.15 vs 129.73 ms — the kernel is not waiting on those loads.
#ifndef G16_BA_ADD_THREADS
#define G16_BA_ADD_THREADS 768   // resident threads per SM the addition kernel is compiled for: 768 (80 registers, 28 bytes of
                                 // spills) against 640 (88 registers): accumulate stage 69.9 vs 70.4 ms per 1024 proofs; 64 pairs per
                                 // thread instead of 32: 70.7 ms (profiles/sweep_r02_add_occupancy.jsonl)
#endif
#ifndef G16_BA_PREFETCH
#define G16_BA_PREFETCH 0
#endif
enum { BA_ADD = 0, BA_DBL = 1, BA_INF = 2, BA_COPYP = 3, BA_COPYQ = 4 };

FD Fp ba_load_fp(const Fp* p) {
#if G16_ASM
    Fp v;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(s), b = __ldg(s + 1);
    v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w; v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
    return v;
#else
    return *p;
#endif
}

// operand `slot` of a level: LEVEL0 -> table point named by refs[slot] (bit 0 = negate, BA_NULL = nothing), else src[slot]
template <bool LEVEL0>
FD bool ba_load_x(const G1Affine* __restrict__ src, const uint32_t* __restrict__ refs, size_t slot, Fp& x, uint32_t& ref) {
    if (LEVEL0) {
        ref = refs[slot];
        if (ref == BA_NULL) { x = Fp::zero(); return false; }
        x = ba_load_fp(&src[ref >> 1].x);
    } else {
        ref = 0;
        x = ba_load_fp(&src[slot].x);
    }
    return true;
}
template <bool LEVEL0>
FD Fp ba_load_y(const G1Affine* __restrict__ src, size_t slot, uint32_t ref) {
    if (LEVEL0) {
        if (ref == BA_NULL) return Fp::zero();
        Fp y = ba_load_fp(&src[ref >> 1].y);
        return (ref & 1u) ? y.neg() : y;
    }
    return ba_load_fp(&src[slot].y);
}

// what the pair (P, Q) needs, and its denominator (never zero). (0,0) is the point at infinity.
FD int ba_classify(const Fp& x1, const Fp& y1, const Fp& x2, const Fp& y2, Fp& d) {
    const bool pinf = x1.is_zero() && y1.is_zero(), qinf = x2.is_zero() && y2.is_zero();
    d = Fp::one();
    if (pinf) return qinf ? BA_INF : BA_COPYQ;
    if (qinf) return BA_COPYP;
    if (x1 == x2) {
        if (y1 == y2 && !y1.is_zero()) { d = y1.dbl(); return BA_DBL; }
        return BA_INF;   // P + (-P)
    }
    d = x2 - x1;
    return BA_ADD;
}
// numerator of the tangent slope, kept out of line: the doubling case is rare (two equal points in one bucket)
#if defined(G16_EMU)
#define G16_BA_NOINLINE
#else
#define G16_BA_NOINLINE __noinline__
#endif
__device__ __host__ G16_BA_NOINLINE static Fp ba_tangent_numerator(const Fp& x) {
    Fp xx = x.sqr();
    return xx.dbl() + xx;
}

// pairs of thread t of block b: q = (b M + j) T + t, j < M  (consecutive threads read consecutive pairs)
template <bool LEVEL0, int T, int M>
__global__ void __launch_bounds__(T)
msm_ba_den_kernel(const G1Affine* __restrict__ src, const uint32_t* __restrict__ refs, const uint32_t* __restrict__ total_slots,
                  int shift, Fp* __restrict__ pre, Fp* __restrict__ tot) {
    const size_t npairs = ((size_t)(*total_slots) >> shift) >> 1;
    const size_t q0 = (size_t)blockIdx.x * M * T + threadIdx.x;
    if ((size_t)blockIdx.x * M * T >= npairs) return;   // whole block past the live length
    Fp c = Fp::one();
#pragma unroll 1
    for (int j = 0; j < M; j++) {
        const size_t q = q0 + (size_t)j * T;
        if (q >= npairs) break;
        Fp x1, x2, d;
        uint32_t r1, r2;
        const bool h1 = ba_load_x<LEVEL0>(src, refs, 2 * q, x1, r1);
        const bool h2 = ba_load_x<LEVEL0>(src, refs, 2 * q + 1, x2, r2);
        d = x2 - x1;
        if (!h1 || !h2 || x1.is_zero() || x2.is_zero() || d.is_zero()) {   // rare: decide on the full points
            Fp y1 = ba_load_y<LEVEL0>(src, 2 * q, r1), y2 = ba_load_y<LEVEL0>(src, 2 * q + 1, r2);
            ba_classify(x1, y1, x2, y2, d);
        }
        c = c * d;
        pre[q] = c;
    }
    tot[(size_t)blockIdx.x * T + threadIdx.x] = c;   // = 1 for a thread without pairs
}

// group size of the inversion kernel for `ntot` live totals: 64 when the machine is full anyway, down to GMIN for the small
// problems (the wire-driven queries), whose time is the latency of one thread's chain (2 G dependent products + one inversion)
FD uint32_t ba_inv_group(size_t ntot, uint32_t gmin, uint32_t gmax) {
    const size_t g = ntot >> 17;   // ~131 k threads fill 148 SMs several times over
    return g < gmin ? gmin : (g > gmax ? gmax : (uint32_t)g);
}
// inverses of the thread totals: thread g owns totals [g G, (g+1) G), G = ba_inv_group(live totals)
template <int T, int M, int GMIN, int GMAX>
__global__ void __launch_bounds__(128)
msm_ba_inv_kernel(const uint32_t* __restrict__ total_slots, int shift, const Fp* __restrict__ tot, Fp* __restrict__ totpre,
                  Fp* __restrict__ totinv) {
    const size_t npairs = ((size_t)(*total_slots) >> shift) >> 1;
    const size_t ntot = ((npairs + (size_t)T * M - 1) / ((size_t)T * M)) * T;   // the den kernel's live threads
    const uint32_t G = ba_inv_group(ntot, GMIN, GMAX);
    const size_t g0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * G;
    if (g0 >= ntot) return;
    const size_t n = ntot - g0 < (size_t)G ? ntot - g0 : (size_t)G;
    Fp c = Fp::one();
#pragma unroll 1
    for (size_t i = 0; i < n; i++) {
        c = c * tot[g0 + i];
        totpre[g0 + i] = c;
    }
    Fp R = c.inv();
#pragma unroll 1
    for (size_t i = n; i-- > 0;) {
        const Fp cprev = i ? totpre[g0 + i - 1] : Fp::one();
        totinv[g0 + i] = cprev * R;
        R = R * tot[g0 + i];
    }
}

// Two-level form of the same step for large levels (millions of thread totals): the binary-Euclid inversion is half of the
// single kernel's issue slots there (one inversion per ~40 totals, every lane on its own divergent path), so the totals are
// first folded into groups of G1 by products alone (fwd), the group totals go through the kernel above (one inversion per
// ~G1 * 4 totals), and a backward walk turns the group inverses into the inverses of the thread totals (bwd). Same values.
template <int T, int M, int G1>
__global__ void __launch_bounds__(128)
msm_ba_inv_fwd_kernel(const uint32_t* __restrict__ total_slots, int shift, const Fp* __restrict__ tot, Fp* __restrict__ totpre,
                      Fp* __restrict__ gtot) {
    const size_t npairs = ((size_t)(*total_slots) >> shift) >> 1;
    const size_t ntot = ((npairs + (size_t)T * M - 1) / ((size_t)T * M)) * T;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t g0 = g * G1;
    if (g0 >= ntot) return;
    const size_t n = ntot - g0 < (size_t)G1 ? ntot - g0 : (size_t)G1;
    Fp c = Fp::one();
#pragma unroll 1
    for (size_t i = 0; i < n; i++) {
        c = c * tot[g0 + i];
        totpre[g0 + i] = c;
    }
    gtot[g] = c;
}
// the kernel above over the group totals: group g owns thread totals [g G1, (g+1) G1)
template <int T, int M, int G1, int GMIN, int GMAX>
__global__ void __launch_bounds__(128)
msm_ba_inv_mid_kernel(const uint32_t* __restrict__ total_slots, int shift, const Fp* __restrict__ gtot, Fp* __restrict__ gpre,
                      Fp* __restrict__ ginv) {
    const size_t npairs = ((size_t)(*total_slots) >> shift) >> 1;
    const size_t ntot = ((npairs + (size_t)T * M - 1) / ((size_t)T * M)) * T;
    const size_t ng = (ntot + G1 - 1) / G1;
    const uint32_t G = ba_inv_group(ng, GMIN, GMAX);
    const size_t g0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * G;
    if (g0 >= ng) return;
    const size_t n = ng - g0 < (size_t)G ? ng - g0 : (size_t)G;
    Fp c = Fp::one();
#pragma unroll 1
    for (size_t i = 0; i < n; i++) {
        c = c * gtot[g0 + i];
        gpre[g0 + i] = c;
    }
    Fp R = c.inv();
#pragma unroll 1
    for (size_t i = n; i-- > 0;) {
        const Fp cprev = i ? gpre[g0 + i - 1] : Fp::one();
        ginv[g0 + i] = cprev * R;
        R = R * gtot[g0 + i];
    }
}
template <int T, int M, int G1>
__global__ void __launch_bounds__(128)
msm_ba_inv_bwd_kernel(const uint32_t* __restrict__ total_slots, int shift, const Fp* __restrict__ tot, const Fp* __restrict__ totpre,
                      const Fp* __restrict__ ginv, Fp* __restrict__ totinv) {
    const size_t npairs = ((size_t)(*total_slots) >> shift) >> 1;
    const size_t ntot = ((npairs + (size_t)T * M - 1) / ((size_t)T * M)) * T;
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t g0 = g * G1;
    if (g0 >= ntot) return;
    const size_t n = ntot - g0 < (size_t)G1 ? ntot - g0 : (size_t)G1;
    Fp R = ginv[g];   // 1 / (product of the group's totals)
#pragma unroll 1
    for (size_t i = n; i-- > 0;) {
        const Fp cprev = i ? totpre[g0 + i - 1] : Fp::one();
        totinv[g0 + i] = cprev * R;
        R = R * tot[g0 + i];
    }
}

template <bool LEVEL0, int T, int M>
__global__ void __launch_bounds__(T, G16_BA_ADD_THREADS / T)
msm_ba_add_kernel(const G1Affine* __restrict__ src, const uint32_t* __restrict__ refs, const uint32_t* __restrict__ total_slots,
                  int shift, const Fp* __restrict__ pre, const Fp* __restrict__ totinv, G1Affine* __restrict__ out) {
    const size_t npairs = ((size_t)(*total_slots) >> shift) >> 1;
    const size_t q0 = (size_t)blockIdx.x * M * T + threadIdx.x;
    if (q0 >= npairs) return;
    int m = (int)((npairs - q0 + T - 1) / T);   // this thread's pairs: j < m
    if (m > M) m = M;
    Fp R = totinv[(size_t)blockIdx.x * T + threadIdx.x];   // 1 / (d_0 ... d_(m-1))
#if G16_BA_PREFETCH
    // the references of pair j - 1 and its running product are requested while pair j is computed: at level 0 the points are
    // two dependent loads away (reference, then table point), and the loop is not unrolled
    uint32_t nr1 = 0, nr2 = 0;
    Fp ncprev = Fp::one();
    {
        const size_t q = q0 + (size_t)(m - 1) * T;
        if (LEVEL0) { nr1 = refs[2 * q]; nr2 = refs[2 * q + 1]; }
        if (m > 1) ncprev = ba_load_fp(&pre[q - T]);
    }
#endif
#pragma unroll 1
    for (int j = m - 1; j >= 0; j--) {
        const size_t q = q0 + (size_t)j * T;
#if G16_BA_PREFETCH
        const Fp cprev = ncprev;
        uint32_t r1 = nr1, r2 = nr2;
        if (j > 0) {
            const size_t qn = q - T;
            if (LEVEL0) { nr1 = refs[2 * qn]; nr2 = refs[2 * qn + 1]; }
            ncprev = j > 1 ? ba_load_fp(&pre[qn - T]) : Fp::one();
        }
        const Fp inv = cprev * R;    // 1 / d_j
        Fp x1, x2, d;
        if (LEVEL0) {
            x1 = r1 == BA_NULL ? Fp::zero() : ba_load_fp(&src[r1 >> 1].x);
            x2 = r2 == BA_NULL ? Fp::zero() : ba_load_fp(&src[r2 >> 1].x);
        } else {
            x1 = ba_load_fp(&src[2 * q].x);
            x2 = ba_load_fp(&src[2 * q + 1].x);
        }
#else
        const Fp cprev = j ? pre[q - T] : Fp::one();
        const Fp inv = cprev * R;    // 1 / d_j
        Fp x1, x2, d;
        uint32_t r1, r2;
        ba_load_x<LEVEL0>(src, refs, 2 * q, x1, r1);
        ba_load_x<LEVEL0>(src, refs, 2 * q + 1, x2, r2);
#endif
        const Fp y1 = ba_load_y<LEVEL0>(src, 2 * q, r1), y2 = ba_load_y<LEVEL0>(src, 2 * q + 1, r2);
        const int kind = ba_classify(x1, y1, x2, y2, d);
        R = R * d;
        G1Affine res;
        if (kind == BA_ADD || kind == BA_DBL) {
            const Fp num = kind == BA_ADD ? (y2 - y1) : ba_tangent_numerator(x1);
            const Fp lam = num * inv;
            const Fp x3 = lam.sqr() - x1 - x2;
            res.x = x3;
            res.y = lam * (x1 - x3) - y1;
        } else if (kind == BA_COPYP) {
            res.x = x1; res.y = y1;
        } else if (kind == BA_COPYQ) {
            res.x = x2; res.y = y2;
        } else {
            res = G1Affine::inf();
        }
        out[q] = res;
    }
}

}  // namespace g16
