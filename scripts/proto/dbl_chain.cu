// Prototype / microbenchmark: dependent chains of G1 XYZZ doublings and additions — what the Horner kernel of a one-shot MSM and the
// scalar multiplications of the proof assembly are made of — run three ways:
//   plain     one thread, the group law of ec.cuh                                          dbl 3.95 us   add 6.32 us   (B200)
//   ilp       one thread, the independent products of a formula interleaved (mul_many)      dbl 4.15 us   add 6.43 us   no gain: a
//             lone warp is bound by its partition's multiplier (~4 cycles per carry-chained wide multiply-add whatever the number
//             of live lanes), not by dependency latency — ptxas had interleaved the chains already
//   team      four warps, one product each per step (team.cuh)                              see gpurun_out / profiles
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I gnark_symmetric_crypto_b200/csrc \
//        scripts/proto/dbl_chain.cu -o gnark_symmetric_crypto_b200/lib/dbl_chain
#include <cstdio>
#include <cstdint>
#include <cstring>
#include "team.cuh"
using namespace g16;

// N INDEPENDENT Montgomery products r[k] = a[k] * b[k] with their rounds interleaved in program order. One product is two carry
// chains of dependent wide multiply-adds, so a thread that runs alone (the latency-bound kernels: Horner over the windows, the
// merge levels and block trees of a small MSM, the scalar multiplications of the proof assembly) waits out the multiplier's
// latency on almost every instruction: ~0.6 us per dependent product. N products side by side give the scheduler 2 N chains to
// alternate between. Same values as operator*, bit for bit (the same rounds, only their order differs). For the throughput
// kernels, whose warps hide the latency, this only costs registers — they keep operator*.
template <class P, int N>
FD void mul_many(Fe<P>* r, const Fe<P>* a, const Fe<P>* b) {
    uint32_t E[N][8], O[N][8], t[N][8], m[N];
#pragma unroll
    for (int k = 0; k < N; k++) {
        mul4(E[k], a[k].l[0], a[k].l[2], a[k].l[4], a[k].l[6], b[k].l[0]);
        mul4(O[k], a[k].l[1], a[k].l[3], a[k].l[5], a[k].l[7], b[k].l[0]);
    }
#pragma unroll
    for (int k = 0; k < N; k++) m[k] = E[k][0] * P::inv();
#pragma unroll
    for (int k = 0; k < N; k++) mad4(O[k], P::mod(1), P::mod(3), P::mod(5), P::mod(7), m[k]);
#pragma unroll
    for (int k = 0; k < N; k++) O[k][7] += mad4(E[k], P::mod(0), P::mod(2), P::mod(4), P::mod(6), m[k]);
#pragma unroll
    for (int i = 1; i < 8; i++) {
#pragma unroll
        for (int k = 0; k < N; k++) mad4_shift(t[k], O[k][0], E[k], a[k].l[1], a[k].l[3], a[k].l[5], a[k].l[7], b[k].l[i]);
#pragma unroll
        for (int k = 0; k < N; k++) t[k][7] += mad4(O[k], a[k].l[0], a[k].l[2], a[k].l[4], a[k].l[6], b[k].l[i]);
#pragma unroll
        for (int k = 0; k < N; k++) m[k] = O[k][0] * P::inv();
#pragma unroll
        for (int k = 0; k < N; k++) mad4(t[k], P::mod(1), P::mod(3), P::mod(5), P::mod(7), m[k]);
#pragma unroll
        for (int k = 0; k < N; k++) t[k][7] += mad4(O[k], P::mod(0), P::mod(2), P::mod(4), P::mod(6), m[k]);
#pragma unroll
        for (int k = 0; k < N; k++) {
#pragma unroll
            for (int j = 0; j < 8; j++) { E[k][j] = O[k][j]; O[k][j] = t[k][j]; }
        }
    }
#pragma unroll
    for (int k = 0; k < N; k++) {
        uint32_t sh[8];
#pragma unroll
        for (int j = 0; j < 7; j++) sh[j] = E[k][j + 1];
        sh[7] = 0;
        add8(r[k].l, O[k], sh);
        r[k].reduce_once();
    }
}


// ---- the same three operations for threads that run alone (G1 only: Fp = Fp). The independent products of a formula go
// through mul_many (field.cuh) side by side: a doubling is 3 dependent steps instead of 9 dependent products, an addition
// 4 instead of 14, a mixed addition 4 instead of 10. Same formulas, same exceptional cases, same values.
FD G1XYZZ dbl_ilp(const G1XYZZ& p) {
const Fp &X = p.X, &Y = p.Y, &ZZ = p.ZZ, &ZZZ = p.ZZZ;
    if (p.is_inf()) return p;
    const Fp U = Y.dbl();
    Fp a[4], b[4], r[4];
    a[0] = U; b[0] = U; a[1] = X; b[1] = X;
    mul_many<FpParams, 2>(r, a, b);
    const Fp V = r[0], M = r[1].dbl() + r[1];
    a[0] = U; b[0] = V; a[1] = X; b[1] = V; a[2] = M; b[2] = M;
    mul_many<FpParams, 3>(r, a, b);
    const Fp W = r[0], S = r[1];
    G1XYZZ o;
    o.X = r[2] - S.dbl();
    a[0] = M; b[0] = S - o.X; a[1] = W; b[1] = Y; a[2] = V; b[2] = ZZ; a[3] = W; b[3] = ZZZ;
    mul_many<FpParams, 4>(r, a, b);
    o.Y = r[0] - r[1];
    o.ZZ = r[2];
    o.ZZZ = r[3];
    return o;
}
FD void add_ilp(G1XYZZ& t, const G1XYZZ& o) {
Fp &X = t.X, &Y = t.Y, &ZZ = t.ZZ, &ZZZ = t.ZZZ;
    if (o.is_inf()) return;
    if (t.is_inf()) { t = o; return; }
    Fp a[4], b[4], r[4];
    a[0] = X; b[0] = o.ZZ; a[1] = o.X; b[1] = ZZ; a[2] = Y; b[2] = o.ZZZ; a[3] = o.Y; b[3] = ZZZ;
    mul_many<FpParams, 4>(r, a, b);
    const Fp U1 = r[0], S1 = r[2];
    const Fp P = r[1] - U1, R = r[3] - S1;
    if (P.is_zero()) {
        if (R.is_zero()) t = dbl_ilp(t);
        else t = G1XYZZ::inf();
        return;
    }
    a[0] = P; b[0] = P; a[1] = R; b[1] = R; a[2] = ZZ; b[2] = o.ZZ; a[3] = ZZZ; b[3] = o.ZZZ;
    mul_many<FpParams, 4>(r, a, b);
    const Fp PP = r[0], RR = r[1], Z2 = r[2], Z3 = r[3];
    a[0] = P; b[0] = PP; a[1] = U1; b[1] = PP; a[2] = Z2; b[2] = PP;
    mul_many<FpParams, 3>(r, a, b);
    const Fp PPP = r[0], Q = r[1];
    ZZ = r[2];
    const Fp X3 = RR - PPP - Q.dbl();
    a[0] = R; b[0] = Q - X3; a[1] = S1; b[1] = PPP; a[2] = Z3; b[2] = PPP;
    mul_many<FpParams, 3>(r, a, b);
    Y = r[0] - r[1];
    X = X3;
    ZZZ = r[2];
}

template <int V>
__global__ void chain_kernel(G1XYZZ* io, int iters) {
    __shared__ Fp sm[TEAM4_SM_ELEMS];
    __shared__ uint32_t flag;
    Team4 T{sm, &flag, (int)(threadIdx.x >> 5), (int)(threadIdx.x & 31), 0};
    if (threadIdx.x == 0) flag = 0;
    __syncthreads();
    if (V < 4 && threadIdx.x) return;
    G1XYZZ acc = io[0], q = io[1];
    for (int i = 0; i < iters; i++) {
        if (V == 0) acc = acc.dbl();
        else if (V == 1) acc = dbl_ilp(acc);
        else if (V == 2) acc.add(q);
        else if (V == 3) add_ilp(acc, q);
        else if (V == 4) acc = team_dbl(T, acc);
        else acc = team_add(T, acc, q, false);
    }
    if (threadIdx.x == 0) io[2 + V] = acc;
}

template <int V>
static float run(G1XYZZ* d, int iters) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    chain_kernel<V><<<1, 128>>>(d, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    chain_kernel<V><<<1, 128>>>(d, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    G1XYZZ h[8];
    memset(h, 0, sizeof(h));
    Fp one = Fp::one();
    Fp two = one + one;
    G1XYZZ g = {one, two, one, one};           // the generator (1, 2)
    h[0] = g.dbl().dbl().dbl();                // 8 G with a non-trivial ZZ
    h[1] = g.dbl();
    h[1].add(g);                               // 3 G
    G1XYZZ* d;
    cudaMalloc(&d, sizeof(h));
    cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    const int iters = 256;
    float t0 = run<0>(d, iters), t1 = run<1>(d, iters), t2 = run<2>(d, iters), t3 = run<3>(d, iters), t4 = run<4>(d, iters), t5 = run<5>(d, iters);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    auto same = [&](int i, int j) { return memcmp(&h[2 + i], &h[2 + j], sizeof(G1XYZZ)) == 0; };
    printf("{\"iters\": %d, \"dbl_us\": %.3f, \"dbl_ilp_us\": %.3f, \"dbl_team_us\": %.3f, \"add_us\": %.3f, \"add_ilp_us\": %.3f, \"add_team_us\": %.3f, "
           "\"same_dbl\": [%d, %d], \"same_add\": [%d, %d], \"err\": \"%s\"}\n",
           iters, 1000 * t0 / iters, 1000 * t1 / iters, 1000 * t4 / iters, 1000 * t2 / iters, 1000 * t3 / iters, 1000 * t5 / iters,
           (int)same(0, 1), (int)same(0, 4), (int)same(2, 3), (int)same(2, 5), cudaGetErrorString(cudaGetLastError()));
    return 0;
}
