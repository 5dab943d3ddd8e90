// Hot translation unit: G1 half of the proof assembly (assemble.cuh) with the Montgomery product inlined.
#include "assemble.cuh"
#include <cstdlib>

namespace g16 {

void launch_fixed_base_table_g1(const G1Affine& base, G1Affine* tab, cudaStream_t st) {
    auto k = fixed_base_table_kernel<G1>;
    G16_LAUNCH(k, 1, FB_WINDOWS, 0, st, false, base, tab);
    G16_CHECK_LAUNCH();
}
// G16_ASSEMBLE_TEAM=0: the one-kernel form (a thread per half-product, everything after the last MSM)
bool assemble_team_enabled() {
    static const int team = [] { const char* v = getenv("G16_ASSEMBLE_TEAM"); return v && *v ? atoi(v) : 1; }();
    return team != 0;
}
// The part of the G1 assembly that only needs the A and B1 query results: Ar, Bs1, -rs delta (phase 1) and the four GLV
// half-products of s*Ar and r*Bs1 by teams of four warps. The caller may run it on its own stream while the Z query is still
// being summed; launch_assemble_finish then only adds the halves to K, Z and -rs delta.
size_t launch_assemble_products(const AssemblyKeys& keys, AssemblyScratch& sc, uint32_t n, const G1XYZZ* mA, const G1XYZZ* mB1,
                                const Fr* rs, cudaStream_t st) {
    sc.Ar.ensure(n); sc.Bs1.ensure(n); sc.nrsd.ensure(n); sc.win_tab.ensure((size_t)n * 4 * 16);   // 15 window multiples + the half-product
    G16_LAUNCH(assemble_phase1_kernel, dim3(n, 3), FB_WINDOWS, 0, st, true, keys, n, mA, mB1, rs, sc.Ar.p, sc.Bs1.p, sc.nrsd.p);
    G16_LAUNCH(assemble_mul_team_kernel, dim3(div_up(n, 32), 4), 128, 0, st, true, n, (const G1XYZZ*)sc.Ar.p, (const G1XYZZ*)sc.Bs1.p, rs,
               sc.win_tab.p);
    G16_CHECK_LAUNCH();
    return 2;
}
size_t launch_assemble_finish(AssemblyScratch& sc, bool with_commitment, uint32_t n, const G1XYZZ* mK, const G1XYZZ* mZ, uint8_t* out,
                              size_t out_stride, cudaStream_t st) {
    G16_LAUNCH(assemble_finish_kernel, div_up(n, 32), dim3(32, 2), 0, st, false, n, with_commitment ? 1 : 0, mK, mZ, (const G1XYZZ*)sc.Ar.p,
               (const G1XYZZ*)sc.nrsd.p, (const G1XYZZ*)sc.win_tab.p, out, out_stride);
    G16_CHECK_LAUNCH();
    return 1;
}
size_t launch_assemble(const AssemblyKeys& keys, AssemblyScratch& sc, bool with_commitment, uint32_t n, const G1XYZZ* mA,
                       const G1XYZZ* mB1, const G1XYZZ* mK, const G1XYZZ* mZ, const Fr* rs, uint8_t* out, size_t out_stride,
                       cudaStream_t st) {
    if (assemble_team_enabled())
        return launch_assemble_products(keys, sc, n, mA, mB1, rs, st) + launch_assemble_finish(sc, with_commitment, n, mK, mZ, out, out_stride, st);
    sc.Ar.ensure(n); sc.Bs1.ensure(n); sc.nrsd.ensure(n); sc.win_tab.ensure((size_t)n * 4 * 16);
    G16_LAUNCH(assemble_phase1_kernel, dim3(n, 3), FB_WINDOWS, 0, st, true, keys, n, mA, mB1, rs, sc.Ar.p, sc.Bs1.p, sc.nrsd.p);
    G16_LAUNCH(assemble_phase2_kernel, div_up(n, 32), dim3(32, 5), 0, st, true, n, with_commitment ? 1 : 0, mK, mZ,
               (const G1XYZZ*)sc.Ar.p, (const G1XYZZ*)sc.Bs1.p, (const G1XYZZ*)sc.nrsd.p, rs, sc.win_tab.p, out, out_stride);
    G16_CHECK_LAUNCH();
    return 2;
}

}  // namespace g16
