// Batch-affine pairwise levels of the G1 bucket accumulation (msm_ba.cuh). Hot TU of its own: the Montgomery product is
// inlined at eight sites of one kernel, and a separate translation unit keeps the build parallel.
#include "msm_ba.cuh"
#include <cstdlib>

namespace g16 {

static const int BA_T = 128;   // threads per block
#ifndef G16_BA_M
#define G16_BA_M 32
#endif
static const int BA_M = G16_BA_M;   // pairs per thread: BA_M additions share one thread total (measured on B200: 8 / 16 / 32 ->
                                    // 155.4 / 150.5 / 148.5 ms per 1024-proof step)
static const int BA_GMIN = 4, BA_GMAX = 64;   // thread totals per inversion: 64 on a full machine, fewer for small problems
static const int BA_G1 = 64;                  // two-level inversion (levels with >= BA_TWO_LEVEL_MIN thread totals): totals per group

// scratch layout (Fp elements): running products of every pair of the largest level, then three arrays of thread totals
size_t msm_ba_scratch_elems(size_t max_slots) {
    const size_t npairs = max_slots >> 1;
    const size_t ntot = ((npairs + (size_t)BA_T * BA_M - 1) / ((size_t)BA_T * BA_M)) * BA_T;
    return npairs + 3 * ntot + 3 * (ntot / BA_G1 + 64) + 64;
}

// K pairwise levels over `max_slots` (upper bound; the live count is *total_slots on the device) slots.
// lvl[l] receives the (slots >> (l+1)) sums of level l.
void msm_ba_levels(const G1Affine* bases, const uint32_t* refs, const uint32_t* total_slots, size_t max_slots, int K,
                   G1Affine* const* lvl, Fp* scratch, cudaStream_t stream) {
    const size_t npairs0 = max_slots >> 1;
    const size_t ntot0 = ((npairs0 + (size_t)BA_T * BA_M - 1) / ((size_t)BA_T * BA_M)) * BA_T;
    Fp* pre = scratch;
    Fp* tot = scratch + npairs0;
    Fp* totpre = tot + ntot0;
    Fp* totinv = totpre + ntot0;
    Fp* gtot = totinv + ntot0;
    Fp* gpre = gtot + (ntot0 / BA_G1 + 64);
    Fp* ginv = gpre + (ntot0 / BA_G1 + 64);
    // G16_BA_INV2=0: always the single inversion kernel; G16_BA_INV2_MIN: thread totals from which a level takes the two-level form
    static const int two_level = [] { const char* v = getenv("G16_BA_INV2"); return v && *v ? atoi(v) : 1; }();
    static const size_t BA_TWO_LEVEL_MIN = [] { const char* v = getenv("G16_BA_INV2_MIN"); return v && *v ? (size_t)atol(v) : ((size_t)1 << 17); }();
    for (int l = 0; l < K; l++) {
        const size_t npairs_max = max_slots >> (l + 1);
        const size_t nblocks = (npairs_max + (size_t)BA_T * BA_M - 1) / ((size_t)BA_T * BA_M);
        if (!nblocks) break;
        const size_t ntot = nblocks * BA_T;
        const unsigned inv_blocks = div_up((ntot + BA_GMIN - 1) / BA_GMIN, 128);   // sized for the smallest group; idle blocks exit
        if (l == 0) {
            auto k1 = msm_ba_den_kernel<true, BA_T, BA_M>;
            G16_LAUNCH(k1, (unsigned)nblocks, BA_T, 0, stream, false, bases, refs, total_slots, 0, pre, tot);
        } else {
            auto k1 = msm_ba_den_kernel<false, BA_T, BA_M>;
            G16_LAUNCH(k1, (unsigned)nblocks, BA_T, 0, stream, false, (const G1Affine*)lvl[l - 1], (const uint32_t*)nullptr, total_slots, l, pre, tot);
        }
        if (two_level && ntot >= BA_TWO_LEVEL_MIN) {
            const size_t ng = (ntot + BA_G1 - 1) / BA_G1;
            auto f = msm_ba_inv_fwd_kernel<BA_T, BA_M, BA_G1>;
            auto m = msm_ba_inv_mid_kernel<BA_T, BA_M, BA_G1, BA_GMIN, BA_GMAX>;
            auto b = msm_ba_inv_bwd_kernel<BA_T, BA_M, BA_G1>;
            G16_LAUNCH(f, div_up(ng, 128), 128, 0, stream, false, total_slots, l, (const Fp*)tot, totpre, gtot);
            G16_LAUNCH(m, div_up((ng + BA_GMIN - 1) / BA_GMIN, 128), 128, 0, stream, false, total_slots, l, (const Fp*)gtot, gpre, ginv);
            G16_LAUNCH(b, div_up(ng, 128), 128, 0, stream, false, total_slots, l, (const Fp*)tot, (const Fp*)totpre, (const Fp*)ginv, totinv);
        } else {
            auto k2 = msm_ba_inv_kernel<BA_T, BA_M, BA_GMIN, BA_GMAX>;
            G16_LAUNCH(k2, inv_blocks, 128, 0, stream, false, total_slots, l, (const Fp*)tot, totpre, totinv);
        }
        if (l == 0) {
            auto k3 = msm_ba_add_kernel<true, BA_T, BA_M>;
            G16_LAUNCH(k3, (unsigned)nblocks, BA_T, 0, stream, false, bases, refs, total_slots, 0, (const Fp*)pre, (const Fp*)totinv, lvl[0]);
        } else {
            auto k3 = msm_ba_add_kernel<false, BA_T, BA_M>;
            G16_LAUNCH(k3, (unsigned)nblocks, BA_T, 0, stream, false, (const G1Affine*)lvl[l - 1], (const uint32_t*)nullptr, total_slots, l,
                       (const Fp*)pre, (const Fp*)totinv, lvl[l]);
        }
    }
    G16_CHECK_LAUNCH();
}

}  // namespace g16
