// Host-visible types of the MSM pipeline (see msm.cuh for the kernels) and the per-group entry points, each compiled in
// its own translation unit: k_msm_g1.cu (hot: Montgomery product inlined) and k_msm_g2.cu (cold).
#pragma once
#include "common.cuh"

namespace g16 {

struct MsmShape {
    int c;          // window bits
    int nwin;       // ceil(254 / c)
    int nbk;        // buckets per window = 2^(c-1)
    int precomp;    // bases = nwin tables, table w holds 2^(c*w) P_i ; all windows share one bucket set
    uint32_t n;     // points per row
    uint32_t rows;
    FD uint32_t buckets_per_row() const { return precomp ? (uint32_t)nbk : (uint32_t)nwin * nbk; }
    FD uint32_t segs_per_row() const { return precomp ? 1u : (uint32_t)nwin; }
};

static const uint32_t MSM_INVALID = 0xFFFFFFFFu;
static const int MSM_TREE_LOG_G = 5, MSM_TREE_G = 1 << MSM_TREE_LOG_G;   // arity of the bucket-reduction tree (throughput)
static const int MSM_TREE_LOG_G_SMALL = 2;                              // arity for small problems (latency)

template <class C>
struct MsmWorkspace {
    typedef typename C::X X;
    DevBuf<uint32_t> counts, tile_sums, total, part_key[2];   // total: [0] sorted slots, [1] entries (= additions), [2] slots >> K
    DevBuf<uint2> entries;
    DevBuf<X> buckets, part_val[2], lvlR[2], lvlV[2], result;
    // batch-affine path (G1, msm_ba.cuh): padded point references, one level buffer per pairwise level, running products
    DevBuf<uint32_t> ba_refs;
    DevBuf<unsigned long long> cursor64;   // placement word per bucket (msm.cuh msm_ba_place)
    DevBuf<typename C::A> ba_lvl[3];
    DevBuf<typename C::F> ba_scratch;
    size_t launches = 0;
    // per run since log_reset(): three words — the number of entries (= bucket additions, the algorithmic work), the number of
    // sorted slots (entries in runs + the padding of the batch-affine path; equal to the entries without it) and the number of
    // entries the XYZZ accumulation walks (group sums + direct leftovers; the entries without the batch-affine path)
    DevBuf<uint32_t> entry_log;
    size_t log_n = 0;
    int last_K = 0;   // batch-affine levels of the last run (0 = XYZZ accumulation only)
    bool no_ba = false;   // set by the caller for queries whose live entries are far below the upper bound (mostly 0 / +-1 scalars)
    void log_reset() { log_n = 0; }
    uint64_t log_sum(cudaStream_t st, int word = 0) {   // synchronises the stream; word 0 = entries, 1 = slots, 2 = XYZZ entries
        if (!log_n) return 0;
        std::vector<uint32_t> h(3 * log_n);
        entry_log.download(h.data(), 3 * log_n, st);
        G16_CUDA(cudaStreamSynchronize(st));
        uint64_t s = 0;
        for (size_t i = 0; i < log_n; i++) s += h[3 * i + word];
        return s;
    }
};

static inline int msm_pick_window(size_t n) {
    // argmin over c of ceil(254/c) * (n + 2^c)  (SURVEY §8d adds_alg), clamped to what the key layout supports
    int best = 4;
    double bc = 1e300;
    for (int c = 4; c <= 22; c++) {
        double cost = (double)((254 + c - 1) / c) * ((double)n + (double)(1u << c));
        if (cost < bc) { bc = cost; best = c; }
    }
    return best;
}
static inline MsmShape msm_make_shape(uint32_t n, uint32_t rows, int c, int precomp) {
    MsmShape s;
    s.c = c;
    s.nwin = (254 + c - 1) / c;
    s.nbk = 1 << (c - 1);
    s.precomp = precomp;
    s.n = n;
    s.rows = rows;
    return s;
}


void msm_run_g1(MsmWorkspace<G1>& ws, const MsmShape& sh, const G1Affine* bases, const Fr* scalars, size_t row_stride,
                size_t elem_stride, const uint32_t* map, int is_mont, cudaStream_t stream, StageTimer* tm = nullptr, int chunk_len = 0);
void msm_run_g2(MsmWorkspace<G2>& ws, const MsmShape& sh, const G2Affine* bases, const Fr* scalars, size_t row_stride,
                size_t elem_stride, const uint32_t* map, int is_mont, cudaStream_t stream, StageTimer* tm = nullptr, int chunk_len = 0);
// table[w*n + i] = 2^(c*w) * P_i
void msm_precompute_g1(const G1Affine* pts, uint32_t n, int nwin, int c, G1Affine* table, cudaStream_t stream);
void msm_precompute_g2(const G2Affine* pts, uint32_t n, int nwin, int c, G2Affine* table, cudaStream_t stream);
void xyzz_add_g1(G1XYZZ* a, const G1XYZZ* b, uint32_t n, cudaStream_t stream);   // a[i] += b[i]
// out[j] = affine(Sum_k w^(-jk) * scale[brev(k)] * Z[brev(k)]), j < n_out (sign flipped if negate); work: 2^lg elements
void group_dft_g1(const G1Affine* Z, uint32_t nZ, int lg, const Fr* scale, int negate, const Fr* tw_inv, G1XYZZ* work,
                  uint32_t n_out, G1Affine* out, cudaStream_t stream);
void xyzz_to_affine_g1(const G1XYZZ* in, uint32_t n, G1Affine* out, cudaStream_t stream);
// batch-affine pairwise levels (k_msm_ba.cu): K rounds of "slot 2q + slot 2q+1 -> slot q", one shared inversion per 1024 pairs
static const int MSM_BA_MAX_LEVELS = 3;
size_t msm_ba_scratch_elems(size_t max_slots);
void msm_ba_levels(const G1Affine* bases, const uint32_t* refs, const uint32_t* total_slots, size_t max_slots, int K,
                   G1Affine* const* lvl, Fp* scratch, cudaStream_t stream);
void xyzz_to_affine_g2(const G2XYZZ* in, uint32_t n, G2Affine* out, cudaStream_t stream);
void xyzz_add_g2(G2XYZZ* a, const G2XYZZ* b, uint32_t n, cudaStream_t stream);   // a[i] += b[i]
// plain per-row sums of point references (msm.cuh msm_sum_rows): entries[j] = (row, index << 1 | negate), sorted by row,
// MSM_INVALID = empty slot; out[row] = the sum (XYZZ)
void msm_sum_rows_g1(MsmWorkspace<G1>& ws, const G1Affine* bases, const uint2* entries, uint32_t n_entries, uint32_t rows,
                     G1XYZZ* out, cudaStream_t stream);
void msm_sum_rows_g2(MsmWorkspace<G2>& ws, const G2Affine* bases, const uint2* entries, uint32_t n_entries, uint32_t rows,
                     G2XYZZ* out, cudaStream_t stream);

// ---- combination tables for wire-driven queries whose scalars are bits (k_bitq.cu)
static const int BITQ_K = 8;                       // wires per binary group: 255 non-empty subset sums; every group owns 2^8 table slots
static const int BITQ_T = 5, BITQ_T_ENTRIES = 243; // wires per ternary group ({0, 1, -1}): 3^5 signed combinations (index in base 3, 2 = -1)
static const uint32_t BITQ_NONE = 0xFFFFFFFFu;     // padding of the last group of a kind
// flags[w]: bit 0 stays set while wire w only holds 0 / 1 (Montgomery), bit 1 while it only holds 0 / 1 / -1, over the `rows`
// witnesses (flags: one word per wire, preset to 3)
void bitq_profile(const Fr* W, size_t wire_stride, uint32_t nb_wires, uint32_t rows, uint32_t* flags, cudaStream_t stream);
// groups [0, groups_bin) are binary (8 points), the others ternary (5 points); 8 slots of grp_pts / grp_wires per group.
// table[g * 256 + idx] = the combination `idx` of the group's points (affine; entry 0 unused)
void bitq_build_g1(const G1Affine* pts, const uint32_t* grp_pts, uint32_t groups, uint32_t groups_bin, G1Affine* table, cudaStream_t stream);
void bitq_build_g2(const G2Affine* pts, const uint32_t* grp_pts, uint32_t groups, uint32_t groups_bin, G2Affine* table, cudaStream_t stream);
// entries[row * groups + g] = (row, table reference of the pattern of group g in witness `row`), MSM_INVALID for the empty
// pattern; *exception is set when a wire of a group holds something its kind of group does not allow
void bitq_entries(const Fr* W, size_t wire_stride, uint32_t rows, const uint32_t* grp_wires, uint32_t groups, uint32_t groups_bin,
                  uint2* entries, uint32_t* exception, cudaStream_t stream);
// out[w * n_sub + j] = table[w * n + sub[j]]  (window tables of the points that stay on the general path)
void bitq_gather_g1(const G1Affine* table, uint32_t n, int nwin, const uint32_t* sub, uint32_t n_sub, G1Affine* out, cudaStream_t stream);
void bitq_gather_g2(const G2Affine* table, uint32_t n, int nwin, const uint32_t* sub, uint32_t n_sub, G2Affine* out, cudaStream_t stream);

}  // namespace g16
