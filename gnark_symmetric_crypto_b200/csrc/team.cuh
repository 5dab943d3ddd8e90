// G1 doubling / addition by a TEAM of four warps, for the kernels that are one long chain of dependent point operations: the
// Horner pass over the windows of a one-shot MSM (msm.cuh msm_combine_team_kernel) and the scalar multiplications s*Ar, r*Bs1
// of the proof assembly (assemble.cuh). Replaces nothing in the reference by itself — it is how a14 / a16 run
// their serial tails; formulas and exceptional cases are those of ec.cuh (EFD dbl-2008-s-1, add-2008-s), so the values are the
// same bit for bit.
//
// Why warps and not lanes: a thread that runs alone issues warp-wide instructions with one live lane, and its scheduler
// partition's multiplier takes ~4 cycles per carry-chained wide multiply-add whatever the number of live lanes — measured on
// B200, a dependent doubling costs 3.95 us (9 products), an addition 6.3 us (14), and interleaving the independent products of
// a formula inside ONE thread (XYZZ::dbl_ilp, scripts/proto/dbl_chain.cu) changes nothing: 4.15 us. The four warps of a
// 128-thread block sit on the SM's four partitions, each with its own multiplier; a formula's independent products go one to
// each warp and meet in shared memory: a doubling is 3 product steps (2 + 3 + 4 products), an addition 4 (4 + 4 + 3 + 3).
// Lane l of every warp of the team works on the same instance l (up to 32 independent chains per team); every warp keeps the
// whole point in registers, only the products travel. One __syncthreads() per step: a block is one team.
#pragma once
#include "ec.cuh"

namespace g16 {

#define TD __device__ __forceinline__   // device only: these synchronise the block

struct Team4 {
    Fp* sm;           // 2 buffers x 4 warps x 32 lanes
    uint32_t* flag;   // one shared word, zero between calls: "a lane of this addition met P = +-Q"
    int w, lane, buf;
};
static const int TEAM4_SM_ELEMS = 2 * 4 * 32;

// every thread of the block calls this the same number of times. Two buffers are enough: a warp can only be one barrier ahead.
TD void team_exchange(Team4& T, const Fp& mine, Fp out[4]) {
    T.sm[(T.buf * 4 + T.w) * 32 + T.lane] = mine;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) out[k] = T.sm[(T.buf * 4 + k) * 32 + T.lane];
    T.buf ^= 1;
}
TD Fp team_sel(int w, const Fp& a0, const Fp& a1, const Fp& a2, const Fp& a3) {
    Fp r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = w == 0 ? a0.l[i] : (w == 1 ? a1.l[i] : (w == 2 ? a2.l[i] : a3.l[i]));
    return r;
}

TD G1XYZZ team_dbl(Team4& T, const G1XYZZ& p) {
    const Fp U = p.Y.dbl();
    Fp o[4], a, b, r = Fp::zero();
    a = team_sel(T.w, U, p.X, p.X, p.X);
    if (T.w < 2) r = a * a;                                // V = U^2 | X^2
    team_exchange(T, r, o);
    const Fp V = o[0], M = o[1].dbl() + o[1];
    a = team_sel(T.w, U, p.X, M, M);
    b = team_sel(T.w, V, V, M, M);
    if (T.w < 3) r = a * b;                                // W = U V | S = X V | M^2
    team_exchange(T, r, o);
    const Fp W = o[0], S = o[1];
    G1XYZZ q;
    q.X = o[2] - S.dbl();
    a = team_sel(T.w, M, W, V, W);
    b = team_sel(T.w, S - q.X, p.Y, p.ZZ, p.ZZZ);
    r = a * b;                                             // M (S - X3) | W Y | V ZZ | W ZZZ
    team_exchange(T, r, o);
    q.Y = o[0] - o[1];
    q.ZZ = o[2];
    q.ZZZ = o[3];
    return p.is_inf() ? p : q;
}

// p + q (skip: the lane keeps p). Lanes whose operands are equal or opposite are counted in the shared flag; only then does the
// whole team take the detour through team_dbl.
TD G1XYZZ team_add(Team4& T, const G1XYZZ& p, const G1XYZZ& q, bool skip) {
    Fp o[4], a, b, r = Fp::zero();
    a = team_sel(T.w, p.X, q.X, p.Y, q.Y);
    b = team_sel(T.w, q.ZZ, p.ZZ, q.ZZZ, p.ZZZ);
    r = a * b;                                             // U1 | U2 | S1 | S2
    team_exchange(T, r, o);
    const Fp U1 = o[0], S1 = o[2];
    const Fp P = o[1] - U1, R = o[3] - S1;
    const bool plain = skip || p.is_inf() || q.is_inf();
    const bool exc = !plain && P.is_zero();
    if (exc) atomicOr(T.flag, 1u);
    a = team_sel(T.w, P, R, p.ZZ, p.ZZZ);
    b = team_sel(T.w, P, R, q.ZZ, q.ZZZ);
    r = a * b;                                             // PP | RR | ZZ1 ZZ2 | ZZZ1 ZZZ2
    team_exchange(T, r, o);
    const bool any_exc = *T.flag != 0u;                    // written before the barrier above, read after it
    const Fp PP = o[0], RR = o[1], Z2 = o[2], Z3 = o[3];
    a = team_sel(T.w, P, U1, Z2, Z2);
    if (T.w < 3) r = a * PP;                               // PPP | Q | ZZ3
    team_exchange(T, r, o);
    const Fp PPP = o[0], Q = o[1];
    G1XYZZ s;
    s.ZZ = o[2];
    s.X = RR - PPP - Q.dbl();
    a = team_sel(T.w, R, S1, Z3, Z3);
    b = team_sel(T.w, Q - s.X, PPP, PPP, PPP);
    if (T.w < 3) r = a * b;                                // R (Q - X3) | S1 PPP | ZZZ3
    team_exchange(T, r, o);
    s.Y = o[0] - o[1];
    s.ZZZ = o[2];
    if (any_exc) {
        __syncthreads();                                   // every thread has read the flag
        if (T.w == 0 && T.lane == 0) *T.flag = 0u;
        const G1XYZZ d = team_dbl(T, p);
        if (exc) s = R.is_zero() ? d : G1XYZZ::inf();
    }
    if (skip || q.is_inf()) return p;
    if (p.is_inf()) return q;
    return s;
}

}  // namespace g16
