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
void xyzz_to_affine_g1(const G1XYZZ* in, uint32_t n, G1Affine* out, cudaStream_t stream) {
    auto k = xyzz_to_affine_kernel<G1>;
    G16_LAUNCH(k, div_up(n, 32), 32, 0, stream, false, in, n, out);
    G16_CHECK_LAUNCH();
}

}  // namespace g16
