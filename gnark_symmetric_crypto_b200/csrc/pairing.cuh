// Optimal ate pairing product check on BN254 for sm_100a: prod_i e(P_i, Q_i) == 1, many independent checks per launch.
// This is the arithmetic under groth16.Verify.
//
// Replaces: gnark-crypto v0.14.0 ecc/bn254/pairing.go (PairingCheck = MillerLoop + FinalExponentiation) as called by
// gnark v0.11.0 backend/groth16/bn254/verify.go (Verify), reached from libraries/verifier/impl/verifiers.go:93-99,139-145.
//
// Formulation (the one the oracle is pinned with, oracle/oracle_groth16.cpp): Fp12 = Fp[w]/(w^12 - 18 w^6 + 82) — twelve
// Fp coefficients, u = w^6 - 9; affine Miller loop on the twist. Final exponentiation: with F = f^((p^2+1) h),
// h = (p^4-p^2+1)/r, the pairing product is 1 iff F lies in Fp6 (its odd coefficients vanish), because
// f^((p^12-1)/r) = conj(F)/F; F is one 254-bit simultaneous exponentiation over g, pi(g), pi^2(g) (pi = Frobenius, a
// coefficient-wise map in this basis) instead of the oracle's plain 2790-bit power. Any correct pairing gives the same
// accept/reject bit.
//
// Mapping:
//   pairing_lines_kernel   thread = one (P, Q) pair: the G2 arithmetic of the loop (102 steps, one Fp2 inversion each) and the
//                          five non-zero Fp coefficients of every line evaluated at P; a chain of dependent products, so
//                          its latency (~25 ms) does not depend on the batch size up to thousands of pairs
//   pairing_check_kernel   block = one check: f <- f^2 * prod(lines) over the steps, then the Fp6-membership test of F. An Fp12 product
//                          is spread over 144 threads (one coefficient product each), the column sums over 23, the
//                          reduction by w^12 = 18 w^6 - 82 over 11: an Fp12 product costs about three dependent Fp products.
#pragma once
#include "common.cuh"
#include "pairing_consts.hpp"

namespace g16 {

static const int PAIRING_THREADS = 160;   // 144 coefficient products + one spare warp; all threads take part in the barriers

struct LineRec {
    Fp c0, c1, c7, c3, c9;   // coefficients of w^0, w^1, w^7, w^3, w^9 ; all other coefficients are zero
};

// ---------------------------------------------------------------------------------------------------- G2 side (one thread)
struct PairingConsts {
    Fp2 frob_x, frob_y;   // xi^((p-1)/3), xi^((p-1)/2), xi = 9 + u
    Fp2 gamma[6];         // gamma[j] = xi^(j (p-1)/6) = w^(j (p-1)): the p-power Frobenius maps e w^j to conj(e) gamma[j] w^j
};
FD Fp fp_small(uint32_t v) {
    Fp x = Fp::zero();
    x.l[0] = v;
    return x.to_mont();
}
__global__ void pairing_consts_kernel(PairingConsts* out) {
    if (threadIdx.x || blockIdx.x) return;
    Fp2 xi = {fp_small(9), Fp::one()};
    uint32_t e3[8], e2[8];
    for (int i = 0; i < 8; i++) e3[i] = PAIRING_PM1_DIV3[i];
    for (int i = 0; i < 8; i++) e2[i] = (FpParams::mod(i) >> 1) | (i < 7 ? (FpParams::mod(i + 1) << 31) : 0u);   // (p-1)/2
    out->frob_x = xi.pow(e3);
    out->frob_y = xi.pow(e2);
    uint32_t e6[8];
    for (int i = 0; i < 8; i++) e6[i] = PAIRING_PM1_DIV6[i];
    Fp2 g1 = xi.pow(e6);
    out->gamma[0] = Fp2::one();
    for (int j = 1; j < 6; j++) out->gamma[j] = out->gamma[j - 1] * g1;
}

// e * w^k for e in Fp2 contributes (e.a0 - 9 e.a1) to the coefficient of w^k and e.a1 to that of w^(k+6)
FD void fp2_at(const Fp2& e, Fp& lo, Fp& hi) {
    Fp t = e.a1.dbl().dbl().dbl() + e.a1;   // 9 a1
    lo = e.a0 - t;
    hi = e.a1;
}
// line through R with slope m (on the twist), evaluated at P
FD LineRec line_eval(const G2Affine& R, const Fp2& m, const G1Affine& P) {
    LineRec l;
    l.c0 = P.y.neg();
    fp2_at(m.mul_fp(P.x), l.c1, l.c7);
    fp2_at(R.y - m * R.x, l.c3, l.c9);
    return l;
}
FD LineRec line_one() {
    LineRec l;
    l.c0 = Fp::one();
    l.c1 = l.c7 = l.c3 = l.c9 = Fp::zero();
    return l;
}

// recs[step * stride + pair]. A pair with P or Q at infinity contributes e = 1: all its lines are the constant 1.
__global__ void __launch_bounds__(64)
pairing_lines_kernel(const G1Affine* __restrict__ Ps, const G2Affine* __restrict__ Qs, uint32_t n_pairs,
                     const PairingConsts* __restrict__ consts, LineRec* __restrict__ recs, size_t stride) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pairs) return;
    const G1Affine P = Ps[i];
    const G2Affine Q = Qs[i];
    LineRec* out = recs + i;
    if (P.is_inf() || Q.is_inf()) {
        for (int s = 0; s < PAIRING_STEPS; s++) out[(size_t)s * stride] = line_one();
        return;
    }
    G2Affine R = Q;
    int s = 0;
    for (int b = 63; b >= 0; b--) {
        {   // doubling step
            Fp2 xx = R.x.sqr();
            Fp2 m = (xx.dbl() + xx) * R.y.dbl().inv();
            out[(size_t)(s++) * stride] = line_eval(R, m, P);
            Fp2 x3 = m.sqr() - R.x.dbl();
            Fp2 y3 = m * (R.x - x3) - R.y;
            R.x = x3; R.y = y3;
        }
        if ((PAIRING_ATE_LOOP_LO >> b) & 1ull) {   // addition step with Q
            Fp2 m = (Q.y - R.y) * (Q.x - R.x).inv();
            out[(size_t)(s++) * stride] = line_eval(R, m, P);
            Fp2 x3 = m.sqr() - R.x - Q.x;
            Fp2 y3 = m * (R.x - x3) - R.y;
            R.x = x3; R.y = y3;
        }
    }
    // Frobenius additions: Q1 = pi(Q), -Q2 = -pi^2(Q)
    G2Affine Q1 = {Q.x.conj() * consts->frob_x, Q.y.conj() * consts->frob_y};
    G2Affine nQ2 = {Q1.x.conj() * consts->frob_x, (Q1.y.conj() * consts->frob_y).neg()};
    {
        Fp2 m = (Q1.y - R.y) * (Q1.x - R.x).inv();
        out[(size_t)(s++) * stride] = line_eval(R, m, P);
        Fp2 x3 = m.sqr() - R.x - Q1.x;
        Fp2 y3 = m * (R.x - x3) - R.y;
        R.x = x3; R.y = y3;
    }
    {
        Fp2 m = (nQ2.y - R.y) * (nQ2.x - R.x).inv();
        out[(size_t)(s++) * stride] = line_eval(R, m, P);
    }
}

// ---------------------------------------------------------------------------------------------------- Fp12, one block
struct F12Shared {
    Fp f[12], g[12], l[12];   // accumulator, scratch, current line / operand
    Fp T[8][12];              // final exponentiation: products of the subsets of {g, pi(g), pi^2(g)} (T[0] unused)
    Fp prod[144];
    Fp col[23];
    Fp hi18[11], hi82[11];
};
// 18 x and 82 x by doubling chains (cheaper than a Montgomery product by a constant)
FD Fp times18(const Fp& x) { Fp x2 = x.dbl(), x16 = x2.dbl().dbl().dbl(); return x16 + x2; }
FD Fp times82(const Fp& x) { Fp x2 = x.dbl(), x16 = x2.dbl().dbl().dbl(), x64 = x16.dbl().dbl(); return x64 + x16 + x2; }

// dst = a * b (dst may alias a or b). Every thread of the block must call.
__device__ __forceinline__ void f12_mul(F12Shared& S, Fp* dst, const Fp* a, const Fp* b) {
    const uint32_t t = threadIdx.x;
    if (t < 144) S.prod[t] = a[t / 12] * b[t % 12];
    __syncthreads();
    if (t < 23) {
        Fp acc = Fp::zero();
        const int lo = (int)t - 11 > 0 ? (int)t - 11 : 0, hi = t < 11 ? (int)t : 11;
        for (int i = lo; i <= hi; i++) acc = acc + S.prod[i * 12 + ((int)t - i)];
        S.col[t] = acc;
    }
    __syncthreads();
    // w^12 = 18 w^6 - 82: coefficients 18..22 fold into 12..16 and 6..10 first, then 12..17 into 6..11 and 0..5
    if (t >= 18 && t < 23) { S.hi18[t - 12] = times18(S.col[t]); S.hi82[t - 12] = times82(S.col[t]); }
    __syncthreads();
    if (t >= 12 && t < 17) S.col[t] = S.col[t] + S.hi18[t - 6];          // T[t] += 18 T[t+6]
    if (t >= 6 && t < 11) S.col[t] = S.col[t] - S.hi82[t];               // T[t] -= 82 T[t+12]
    __syncthreads();
    if (t >= 12 && t < 18) { S.hi18[t - 12] = times18(S.col[t]); S.hi82[t - 12] = times82(S.col[t]); }
    __syncthreads();
    if (t >= 6 && t < 12) dst[t] = S.col[t] + S.hi18[t - 6];           // c[t] = T[t] + 18 T[t+6]
    if (t < 6) dst[t] = S.col[t] - S.hi82[t];                          // c[t] = T[t] - 82 T[t+12]
    __syncthreads();
}

// dst = src^p (dst may alias src). In the tower view src = sum_{j<6} e_j w^j with e_j = (c_j + 9 c_{j+6}) + c_{j+6} u in Fp2, and
// (e_j w^j)^p = conj(e_j) gamma_j w^j. Every thread of the block must call.
__device__ __forceinline__ void f12_frobenius(Fp* dst, const Fp* src, const PairingConsts* __restrict__ consts) {
    const uint32_t t = threadIdx.x;
    Fp lo, hi;
    if (t < 6) {
        Fp b = src[t + 6];
        Fp a = src[t] + (b.dbl().dbl().dbl() + b);   // c_j + 9 c_{j+6}
        Fp2 e = Fp2{a, b.neg()} * consts->gamma[t];
        fp2_at(e, lo, hi);
    }
    __syncthreads();
    if (t < 6) { dst[t] = lo; dst[t + 6] = hi; }
    __syncthreads();
}

// ok[check] = 1 iff prod over its pairs of e(P, Q) == 1. Pair j of check c = pair index c * pairs_per_check + j of `recs`.
__global__ void __launch_bounds__(PAIRING_THREADS)
pairing_check_kernel(const LineRec* __restrict__ recs, size_t stride, uint32_t pairs_per_check,
                     const PairingConsts* __restrict__ consts, uint8_t* __restrict__ ok) {
    __shared__ F12Shared S;
    const uint32_t t = threadIdx.x;
    const uint32_t chk = blockIdx.x;
    const LineRec* base = recs + (size_t)chk * pairs_per_check;
    if (t < 12) S.f[t] = t == 0 ? Fp::one() : Fp::zero();
    __syncthreads();
    auto mul_lines = [&](int step) {
        for (uint32_t j = 0; j < pairs_per_check; j++) {
            if (t < 12) {
                const LineRec* r = base + (size_t)step * stride + j;
                Fp v = Fp::zero();
                if (t == 0) v = r->c0;
                else if (t == 1) v = r->c1;
                else if (t == 7) v = r->c7;
                else if (t == 3) v = r->c3;
                else if (t == 9) v = r->c9;
                S.l[t] = v;
            }
            __syncthreads();
            f12_mul(S, S.f, S.f, S.l);
        }
    };
    int step = 0;
    for (int b = 63; b >= 0; b--) {
        f12_mul(S, S.f, S.f, S.f);
        mul_lines(step++);
        if ((PAIRING_ATE_LOOP_LO >> b) & 1ull) mul_lines(step++);
    }
    mul_lines(step++);
    mul_lines(step++);
    // Final exponentiation, (p^12 - 1)/r = (p^6 - 1)(p^2 + 1) h. The p^6-power Frobenius is the conjugation w -> -w, so
    // f^((p^12-1)/r) = conj(F) / F with F = f^((p^2+1) h): the product of pairings is 1 iff F lies in Fp6, i.e. iff its odd
    // coefficients vanish — no Fp12 inversion is needed. h = H0 + H1 p + H2 p^2 + p^3 (pairing_consts.hpp), hence
    // F = g^H0 pi(g)^H1 pi^2(g)^H2 pi^3(g) with g = pi^2(f) f: one 254-bit simultaneous exponentiation over the three bases.
    f12_frobenius(S.g, S.f, consts);
    f12_frobenius(S.g, S.g, consts);
    f12_mul(S, S.T[1], S.g, S.f);                 // g
    f12_frobenius(S.T[2], S.T[1], consts);        // pi(g)
    f12_frobenius(S.T[4], S.T[2], consts);        // pi^2(g)
    f12_frobenius(S.g, S.T[4], consts);           // pi^3(g), multiplied in at the end
    f12_mul(S, S.T[3], S.T[1], S.T[2]);
    f12_mul(S, S.T[5], S.T[1], S.T[4]);
    f12_mul(S, S.T[6], S.T[2], S.T[4]);
    f12_mul(S, S.T[7], S.T[3], S.T[4]);
    if (t < 12) S.f[t] = t == 0 ? Fp::one() : Fp::zero();
    __syncthreads();
    for (int b = 253; b >= 0; b--) {
        f12_mul(S, S.f, S.f, S.f);
        const uint32_t m = ((PAIRING_HARD_DIGITS[0][b >> 5] >> (b & 31)) & 1u) | (((PAIRING_HARD_DIGITS[1][b >> 5] >> (b & 31)) & 1u) << 1) |
                           (((PAIRING_HARD_DIGITS[2][b >> 5] >> (b & 31)) & 1u) << 2);
        if (m) f12_mul(S, S.f, S.f, S.T[m]);
    }
    f12_mul(S, S.f, S.f, S.g);
    if (t == 0) {
        // F == 0 also has vanishing odd coefficients, but it only arises from a degenerate Miller step on unvalidated input
        // (a zero line: Fe::inv(0) = 0) and is never a product of pairings: rejected.
        bool in_fp6 = true, nonzero = false;
        for (int i = 1; i < 12; i += 2) in_fp6 = in_fp6 && S.f[i].is_zero();
        for (int i = 0; i < 12; i += 2) nonzero = nonzero || !S.f[i].is_zero();
        ok[chk] = (in_fp6 && nonzero) ? 1 : 0;
    }
}

}  // namespace g16
