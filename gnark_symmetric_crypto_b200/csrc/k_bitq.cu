// Combination tables for the wire-driven queries (A, B1, K on G1, B on G2) of a bit-level circuit. Cold translation unit.
//
// Replaces: the MultiExp calls of gnark v0.11.0 backend/groth16/bn254/prove.go:197-260 over
// pk.G1.A, pk.G1.B, pk.G1.K and pk.G2.B (reached from libraries/prover/impl/provers.go:148) for the wires of the witness
// that only ever hold 0 or 1 — all but a few hundred of the 24 k wires of the ChaCha20-V3 circuit.
//
// Sum_i w_i P_i with w_i in {0, 1} is a subset sum: the bit points of a query are cut into groups of 8, every group gets the
// table of its 255 non-empty subset sums (built once, affine), and a witness then costs ONE mixed addition per group instead
// of one per set bit (~4 on average) — no digits, no histogram, no sort, no bucket tree: the 8-bit pattern of a group is the
// table index, the sum of a request's table points is its result. Wires that hold 0, 1 or -1 (the other half of the ChaCha
// circuit: 10.7 k of its 23.3 k wires take all three values, 12.5 k only 0 / 1, none anything else) form TERNARY groups of 5
// with the 242 signed combinations (index in base 3, digit 2 = -1: the scalar r - 1 times P is -P, the same group element). Which wires are bits is LEARNED from the first witnesses
// (g16_ctx.cuh) and CHECKED on every witness: a wire of a group that holds anything else raises the exception flag and the
// context proves that batch again on the general path (and stops using the tables). The points that are not bits stay on the
// general Pippenger path (a sub-query over the gathered window tables), and the two partial results are added.
#include "ec.cuh"
#include "msm_types.hpp"

namespace g16 {

static __global__ void bitq_profile_kernel(const Fr* __restrict__ W, size_t wire_stride, uint32_t nb_wires, uint32_t rows,
                                           uint32_t* __restrict__ flags) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)nb_wires * rows) return;
    const uint32_t w = (uint32_t)(gid / rows), row = (uint32_t)(gid % rows);   // consecutive threads: one wire of consecutive witnesses
    const Fr v = W[(size_t)w * wire_stride + row];
    if (v.is_zero() || v == Fr::one()) return;
    atomicAnd(&flags[w], v == Fr::modulus_minus(Fr::one()) ? 2u : 0u);   // bit 0: only {0, 1} so far; bit 1: only {0, 1, -1}
}
void bitq_profile(const Fr* W, size_t wire_stride, uint32_t nb_wires, uint32_t rows, uint32_t* flags, cudaStream_t stream) {
    G16_LAUNCH(bitq_profile_kernel, div_up((size_t)nb_wires * rows, 256), 256, 0, stream, false, W, wire_stride, nb_wires, rows, flags);
    G16_CHECK_LAUNCH();
}

template <class C>
static __global__ void bitq_build_kernel(const typename C::A* __restrict__ pts, const uint32_t* __restrict__ grp_pts, uint32_t groups,
                                         uint32_t groups_bin, typename C::A* __restrict__ table) {
    typedef typename C::X X;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)groups << BITQ_K) return;
    const uint32_t g = (uint32_t)(gid >> BITQ_K), idx = (uint32_t)(gid & ((1u << BITQ_K) - 1u));
    X acc = X::inf();
    if (g < groups_bin) {
        for (int j = 0; j < BITQ_K; j++) {
            if (!((idx >> j) & 1u)) continue;
            const uint32_t pi = grp_pts[(size_t)g * BITQ_K + j];
            if (pi == BITQ_NONE) continue;
            acc.madd(pts[pi], false);
        }
    } else if (idx < BITQ_T_ENTRIES) {   // ternary group: digit j of idx in base 3 is the coefficient of point j (2 = -1)
        uint32_t rem = idx;
        for (int j = 0; j < BITQ_T; j++) {
            const uint32_t d = rem % 3u;
            rem /= 3u;
            if (!d) continue;
            const uint32_t pi = grp_pts[(size_t)g * BITQ_K + j];
            if (pi == BITQ_NONE) continue;
            acc.madd(pts[pi], d == 2u);
        }
    }
    table[gid] = acc.to_affine();
}
void bitq_build_g1(const G1Affine* pts, const uint32_t* grp_pts, uint32_t groups, uint32_t groups_bin, G1Affine* table, cudaStream_t stream) {
    auto k = bitq_build_kernel<G1>;
    G16_LAUNCH(k, div_up((size_t)groups << BITQ_K, 64), 64, 0, stream, false, pts, grp_pts, groups, groups_bin, table);
    G16_CHECK_LAUNCH();
}
void bitq_build_g2(const G2Affine* pts, const uint32_t* grp_pts, uint32_t groups, uint32_t groups_bin, G2Affine* table, cudaStream_t stream) {
    auto k = bitq_build_kernel<G2>;
    G16_LAUNCH(k, div_up((size_t)groups << BITQ_K, 64), 64, 0, stream, false, pts, grp_pts, groups, groups_bin, table);
    G16_CHECK_LAUNCH();
}

static __global__ void bitq_entries_kernel(const Fr* __restrict__ W, size_t wire_stride, uint32_t rows,
                                           const uint32_t* __restrict__ grp_wires, uint32_t groups, uint32_t groups_bin,
                                           uint2* __restrict__ entries, uint32_t* __restrict__ exception) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)groups * rows) return;
    const uint32_t g = (uint32_t)(gid / rows), row = (uint32_t)(gid % rows);   // a warp reads one wire of 32 witnesses: 1 KB contiguous
    uint32_t idx = 0;
    bool bad = false;
    const Fr one = Fr::one();
    if (g < groups_bin) {
        for (int j = 0; j < BITQ_K; j++) {
            const uint32_t w = grp_wires[(size_t)g * BITQ_K + j];
            if (w == BITQ_NONE) continue;
            const Fr v = W[(size_t)w * wire_stride + row];
            if (v.is_zero()) continue;
            if (v == one) idx |= 1u << j;
            else bad = true;
        }
    } else {   // ternary group: five wires in {0, 1, -1}, index in base 3 (2 = -1)
        const Fr mone = Fr::modulus_minus(one);
        uint32_t p3 = 1;
        for (int j = 0; j < BITQ_T; j++, p3 *= 3u) {
            const uint32_t w = grp_wires[(size_t)g * BITQ_K + j];
            if (w == BITQ_NONE) continue;
            const Fr v = W[(size_t)w * wire_stride + row];
            if (v.is_zero()) continue;
            if (v == one) idx += p3;
            else if (v == mone) idx += 2u * p3;
            else bad = true;
        }
    }
    entries[(size_t)row * groups + g] = make_uint2(row, idx ? (((g << BITQ_K) | idx) << 1) : MSM_INVALID);
    if (bad) *exception = 1u;
}
void bitq_entries(const Fr* W, size_t wire_stride, uint32_t rows, const uint32_t* grp_wires, uint32_t groups, uint32_t groups_bin,
                  uint2* entries, uint32_t* exception, cudaStream_t stream) {
    G16_LAUNCH(bitq_entries_kernel, div_up((size_t)groups * rows, 256), 256, 0, stream, false, W, wire_stride, rows, grp_wires, groups,
               groups_bin, entries, exception);
    G16_CHECK_LAUNCH();
}

template <class A>
static __global__ void bitq_gather_kernel(const A* __restrict__ table, uint32_t n, int nwin, const uint32_t* __restrict__ sub,
                                          uint32_t n_sub, A* __restrict__ out) {
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)n_sub * nwin) return;
    const uint32_t w = (uint32_t)(gid / n_sub), j = (uint32_t)(gid % n_sub);
    out[gid] = table[(size_t)w * n + sub[j]];
}
void bitq_gather_g1(const G1Affine* table, uint32_t n, int nwin, const uint32_t* sub, uint32_t n_sub, G1Affine* out, cudaStream_t stream) {
    if (!n_sub) return;
    auto k = bitq_gather_kernel<G1Affine>;
    G16_LAUNCH(k, div_up((size_t)n_sub * nwin, 128), 128, 0, stream, false, table, n, nwin, sub, n_sub, out);
    G16_CHECK_LAUNCH();
}
void bitq_gather_g2(const G2Affine* table, uint32_t n, int nwin, const uint32_t* sub, uint32_t n_sub, G2Affine* out, cudaStream_t stream) {
    if (!n_sub) return;
    auto k = bitq_gather_kernel<G2Affine>;
    G16_LAUNCH(k, div_up((size_t)n_sub * nwin, 128), 128, 0, stream, false, table, n, nwin, sub, n_sub, out);
    G16_CHECK_LAUNCH();
}

}  // namespace g16
