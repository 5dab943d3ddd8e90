// G1 instantiation of the MSM pipeline (msm.cuh). Hot TU: the field product is inlined into the bucket-accumulation kernel.
#include "msm.cuh"

namespace g16 {

void msm_run_g1(MsmWorkspace<G1>& ws, const MsmShape& sh, const G1Affine* bases, const Fr* scalars, size_t row_stride,
                size_t elem_stride, const uint32_t* map, int is_mont, cudaStream_t stream, StageTimer* tm, int chunk_len) {
    msm_run<G1>(ws, sh, bases, scalars, row_stride, elem_stride, map, is_mont, stream, tm, chunk_len);
}
void msm_precompute_g1(const G1Affine* pts, uint32_t n, int nwin, int c, G1Affine* table, cudaStream_t stream) {
    auto k = msm_precompute_kernel<G1>;
    G16_LAUNCH(k, div_up(n, 64), 64, 0, stream, false, pts, n, nwin, c, table);
    G16_CHECK_LAUNCH();
}
// a[i] += b[i] (the two partial results of the evaluation-basis Z query)
static __global__ void xyzz_add_kernel(G1XYZZ* __restrict__ a, const G1XYZZ* __restrict__ b, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1XYZZ v = a[i];
    v.add(b[i]);
    a[i] = v;
}
void xyzz_add_g1(G1XYZZ* a, const G1XYZZ* b, uint32_t n, cudaStream_t stream) {
    G16_LAUNCH(xyzz_add_kernel, div_up(n, 32), 32, 0, stream, false, a, b, n);
    G16_CHECK_LAUNCH();
}
// ---- DFT over group elements: the evaluation-basis tables of the Z query (g16_ctx.cuh ctx_build_eval_tables).
// With h_p = scale[p] * Sum_j d_j w^(-j brev(p)) (the DIF inverse transform of compute_h_run, p = bit-reversed position):
//   Sum_p h_p Z_p = Sum_j d_j Q_j,   Q_j = Sum_k w^(-jk) S_k,   S_k = scale[brev(k)] * Z[brev(k)],
// i.e. Q is the DFT (root w^-1) of the pre-scaled key points: n scalar products, then log2(n) stages of n/2 butterflies
// (X, Y) -> (X + Y, w^-e (X - Y)), one 254-bit scalar product each. Output of the DIF stages is bit-reversed.
static __global__ void gdft_scale_kernel(const G1Affine* __restrict__ Z, uint32_t nZ, int lg, const Fr* __restrict__ scale, int negate,
                                         G1XYZZ* __restrict__ out) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (1u << lg)) return;
    uint32_t p = __brev(k) >> (32 - lg);
    G1XYZZ r = G1XYZZ::inf();
    if (p < nZ) {
        Fr s = scale[p].from_mont();
        r = scalar_mul(G1XYZZ::from_affine(Z[p]), s);
        if (negate) r = r.neg();
    }
    out[k] = r;
}
static __global__ void gdft_stage_kernel(G1XYZZ* __restrict__ data, int lg, int s, const Fr* __restrict__ tw_inv) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = 1u << lg;
    if (t >= n / 2) return;
    uint32_t half = n >> (s + 1);
    uint32_t off = t & (half - 1);
    uint32_t i = ((t - off) << 1) + off, j = i + half;
    G1XYZZ X = data[i], Y = data[j];
    G1XYZZ sum = X, dif = X;
    sum.add(Y);
    dif.add(Y.neg());
    uint32_t e = off << s;
    if (e) {
        Fr w = tw_inv[e].from_mont();
        dif = scalar_mul(dif, w);
    }
    data[i] = sum;
    data[j] = dif;
}
static __global__ void gdft_finish_kernel(const G1XYZZ* __restrict__ data, int lg, uint32_t n_out, G1Affine* __restrict__ out) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    out[j] = data[__brev(j) >> (32 - lg)].to_affine();
}
void group_dft_g1(const G1Affine* Z, uint32_t nZ, int lg, const Fr* scale, int negate, const Fr* tw_inv, G1XYZZ* work,
                  uint32_t n_out, G1Affine* out, cudaStream_t stream) {
    const uint32_t n = 1u << lg;
    G16_LAUNCH(gdft_scale_kernel, div_up(n, 32), 32, 0, stream, false, Z, nZ, lg, scale, negate, work);
    for (int s = 0; s < lg; s++)
        G16_LAUNCH(gdft_stage_kernel, div_up(n / 2, 32), 32, 0, stream, false, work, lg, s, tw_inv);
    G16_LAUNCH(gdft_finish_kernel, div_up(n_out, 32), 32, 0, stream, false, (const G1XYZZ*)work, lg, n_out, out);
    G16_CHECK_LAUNCH();
}
void msm_sum_rows_g1(MsmWorkspace<G1>& ws, const G1Affine* bases, const uint2* entries, uint32_t n_entries, uint32_t rows,
                     G1XYZZ* out, cudaStream_t stream) {
    msm_sum_rows<G1>(ws, bases, entries, n_entries, rows, out, stream);
}
void xyzz_to_affine_g1(const G1XYZZ* in, uint32_t n, G1Affine* out, cudaStream_t stream) {
    auto k = xyzz_to_affine_kernel<G1>;
    G16_LAUNCH(k, div_up(n, 32), 32, 0, stream, false, in, n, out);
    G16_CHECK_LAUNCH();
}

}  // namespace g16
