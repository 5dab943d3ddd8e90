// G2 instantiation of the MSM pipeline (msm.cuh). Cold TU (built with -DG16_COLD): G2 is 6k additions per ChaCha proof, so code size wins over inlining.
#include "msm.cuh"

namespace g16 {

void msm_run_g2(MsmWorkspace<G2>& ws, const MsmShape& sh, const G2Affine* bases, const Fr* scalars, size_t row_stride,
                size_t elem_stride, const uint32_t* map, int is_mont, cudaStream_t stream, StageTimer* tm, int chunk_len) {
    msm_run<G2>(ws, sh, bases, scalars, row_stride, elem_stride, map, is_mont, stream, tm, chunk_len);
}
void msm_precompute_g2(const G2Affine* pts, uint32_t n, int nwin, int c, G2Affine* table, cudaStream_t stream) {
    auto k = msm_precompute_kernel<G2>;
    G16_LAUNCH(k, div_up(n, 64), 64, 0, stream, false, pts, n, nwin, c, table);
    G16_CHECK_LAUNCH();
}
void msm_sum_rows_g2(MsmWorkspace<G2>& ws, const G2Affine* bases, const uint2* entries, uint32_t n_entries, uint32_t rows,
                     G2XYZZ* out, cudaStream_t stream) {
    msm_sum_rows<G2>(ws, bases, entries, n_entries, rows, out, stream);
}
static __global__ void xyzz_add_g2_kernel(G2XYZZ* __restrict__ a, const G2XYZZ* __restrict__ b, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G2XYZZ v = a[i];
    v.add(b[i]);
    a[i] = v;
}
void xyzz_add_g2(G2XYZZ* a, const G2XYZZ* b, uint32_t n, cudaStream_t stream) {
    G16_LAUNCH(xyzz_add_g2_kernel, div_up(n, 32), 32, 0, stream, false, a, b, n);
    G16_CHECK_LAUNCH();
}
void xyzz_to_affine_g2(const G2XYZZ* in, uint32_t n, G2Affine* out, cudaStream_t stream) {
    auto k = xyzz_to_affine_kernel<G2>;
    G16_LAUNCH(k, div_up(n, 32), 32, 0, stream, false, in, n, out);
    G16_CHECK_LAUNCH();
}

}  // namespace g16
