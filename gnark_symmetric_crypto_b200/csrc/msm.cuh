// Pippenger multi-scalar multiplication for BN254 G1 / G2 on sm_100a.
//
// Replaces: gnark-crypto v0.14.0 ecc/bn254/multiexp.go:106-325 ((*G1Jac).MultiExp, _innerMsmG1,
// partitionScalars :670-792, msmReduceChunkG1Affine), multiexp_affine.go:35-176, :363-500 (G2) — the five MSMs of
// groth16 Prove (gnark backend/groth16/bn254/prove.go:197-295) reached from libraries/prover/impl/provers.go:148,216.
//
// Pipeline (all rows of a batch share the bases; row = one independent scalar vector, i.e. one proof):
//   1 digits+count   signed c-bit windows of sign-normalised scalars (s > r/2 -> r-s with the point negated, so the
//                    {0,1,-1} wires of the ChaCha circuit cost one addition each); histogram of bucket ids
//   2 scan           exclusive prefix sum of the histogram
//   3 scatter        counting sort of (bucket id, point ref) pairs = single-pass radix sort on the full key
//   4 accumulate     every thread owns L consecutive SORTED entries: perfectly balanced, no atomics; complete buckets
//                    are written directly, the two open runs at the chunk edges go to head/tail partials
//   5 merge          partials of a bucket that spans several chunks are summed
//   6 reduce         sum_k k*B_k per (row, window) as a 32-ary tree of running sums
//   7 combine        Horner over windows (not needed with precomputed 2^(cw) P tables: one bucket set per row)
// Points are read with 128-bit loads (64 B / 128 B affine points are 16-byte aligned).
#pragma once
#include "msm_types.hpp"
#include "team.cuh"
#include <cstdlib>
#include <type_traits>

namespace g16 {

// ------------------------------------------------------------------------------------------------ scalar recoding
// canonical, sign-normalised scalar -> signed digits. Returns the digit of window w in [-2^(c-1), 2^(c-1)].
struct Recoded {
    uint32_t s[9];
    uint32_t neg;
};

FD Recoded recode_value(Fr v, int is_mont) {
    if (is_mont) v = v.from_mont();
    // s > (r-1)/2  ->  s = r - s, negate the point
    uint32_t h[8], t[8];
    for (int i = 0; i < 8; i++) h[i] = (FrParams::mod(i) >> 1) | (i < 7 ? (FrParams::mod(i + 1) << 31) : 0u);
    uint32_t big = sub8(t, h, v.l);   // borrow <=> v > h
    Recoded r;
    r.neg = big ? 1u : 0u;
    if (big) v = Fr::modulus_minus(v);
    for (int i = 0; i < 8; i++) r.s[i] = v.l[i];
    r.s[8] = 0;
    return r;
}

// iterate the windows in order; carry must start at 0
FD int recode_digit(const Recoded& r, int w, int c, int& carry) {
    int bit = w * c;
    int limb = bit >> 5, sh = bit & 31;
    uint64_t v = r.s[limb];
    if (limb < 8) v |= (uint64_t)r.s[limb + 1] << 32;
    int d = (int)((v >> sh) & ((1u << c) - 1u)) + carry;
    if (d > (1 << (c - 1))) { d -= (1 << c); carry = 1; } else carry = 0;
    return d;
}

// One thread per (row, point). pass 0: histogram ; pass 1: scatter into the sorted entry array ; pass 2: scatter for the
// batch-affine path (msm_ba.cuh): runs are padded to multiples of 2^ba_shift slots, a slot holds only the point reference
// (refs[pos]), and the first slot of every group of 2^ba_shift writes the group's key (bucket id, group index) to `entries`.
// scalar of (row r, point i) = scalars[r*row_stride + e*elem_stride], e = map ? map[i] : i. Consecutive threads walk the
// unit-stride dimension (points for row-major h vectors, rows for the wire-major witness array).
static const uint32_t MSM_REF_TABLE = 0x80000000u;   // entry reference into the base table (a direct leftover), not a level sum
// Pass 2 of the batch-affine path places every entry through a 64-bit cursor per bucket (built by scan_tiles):
// cursor word = slot offset of the bucket's run (high 32 bits) | direct leftovers of the bucket (bits 29..31) | entries placed
// so far (bits 0..28). The first `lc` entries of a bucket skip the batch-affine levels: they go straight into the accumulate
// list as references into the base table; the others fill the run, whose length is a multiple of 2^ba_shift.
FD void msm_ba_place(unsigned long long word, uint32_t b, uint32_t ref_neg, int ba_shift, const uint32_t* __restrict__ loff,
                     uint2* __restrict__ entries, uint32_t* __restrict__ refs) {
    const uint32_t soff = (uint32_t)(word >> 32), low = (uint32_t)word;
    const uint32_t lc = low >> 29, j = low & 0x1FFFFFFFu;
    const uint32_t ebase = (soff >> ba_shift) + loff[b];   // first accumulate entry of this bucket
    if (j < lc) {
        entries[ebase + j] = make_uint2(b, MSM_REF_TABLE | ref_neg);
    } else {
        const uint32_t k = j - lc, pos = soff + k;
        refs[pos] = ref_neg;
        if ((k & ((1u << ba_shift) - 1u)) == 0) entries[ebase + lc + (k >> ba_shift)] = make_uint2(b, (pos >> ba_shift) << 1);
    }
}
static __global__ void msm_digits_kernel(MsmShape sh, const Fr* __restrict__ scalars, size_t row_stride, size_t elem_stride,
                                  const uint32_t* __restrict__ map, int is_mont, int pass, uint32_t* __restrict__ counts,
                                  uint2* __restrict__ entries, uint32_t* __restrict__ refs = nullptr, int ba_shift = 0,
                                  unsigned long long* __restrict__ cursor64 = nullptr) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)sh.n * sh.rows) return;
    uint32_t row, i;
    if (elem_stride == 1) { row = (uint32_t)(gid / sh.n); i = (uint32_t)(gid % sh.n); }
    else { row = (uint32_t)(gid % sh.rows); i = (uint32_t)(gid / sh.rows); }
    size_t sidx = (size_t)row * row_stride + (size_t)(map ? map[i] : i) * elem_stride;
    Fr raw = scalars[sidx];
    if (raw.is_zero()) return;
    // Montgomery +1 / -1 need no conversion: one digit of magnitude 1 in window 0. The wire-driven queries of a bit-level
    // circuit meet almost nothing else, and their scalars are wire-major (a warp = one wire of 32 proofs), so whole warps take
    // this path: the from_mont products were most of the time of those launches (14 M threads per query and pass).
    int unit = 0;
    if (is_mont) unit = raw == Fr::one() ? 1 : (raw == Fr::modulus_minus(Fr::one()) ? 2 : 0);
    Recoded r;
    if (unit) {
        for (int k = 0; k < 9; k++) r.s[k] = k == 0 ? 1u : 0u;
        r.neg = unit == 2 ? 1u : 0u;
    } else {
        r = recode_value(raw, is_mont);
    }
    const int nwin = unit ? 1 : sh.nwin;
    int carry = 0;
    uint32_t base = row * sh.buckets_per_row();
    for (int w = 0; w < nwin; w++) {
        int d = recode_digit(r, w, sh.c, carry);
        if (d == 0) continue;
        uint32_t neg = r.neg ^ (d < 0 ? 1u : 0u);
        uint32_t mag = (uint32_t)(d < 0 ? -d : d);
        uint32_t b = base + (sh.precomp ? 0u : (uint32_t)w * sh.nbk) + (mag - 1u);
        if (pass == 0) {
            atomicAdd(&counts[b], 1u);
        } else {
            uint32_t ref = sh.precomp ? (uint32_t)w * sh.n + i : i;
            if (pass == 2 && cursor64) {   // counts holds the leftover offsets of the buckets here
                msm_ba_place(atomicAdd(&cursor64[b], 1ull), b, (ref << 1) | neg, ba_shift, counts, entries, refs);
                continue;
            }
            uint32_t pos = atomicAdd(&counts[b], 1u);   // counts holds the running write cursor (= exclusive scan)
            if (pass == 1) {
                entries[pos] = make_uint2(b, (ref << 1) | neg);
            } else {
                refs[pos] = (ref << 1) | neg;
                if ((pos & ((1u << ba_shift) - 1u)) == 0) entries[pos >> ba_shift] = make_uint2(b, (pos >> ba_shift) << 1);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ per-row sort in shared memory
// The same counting sort for batches whose rows share one bucket set (fixed-base tables) of at most 2^14 buckets: ONE block
// per row keeps the row's histogram / write cursors in shared memory, so the two passes cost shared-memory atomics instead of
// one L2 atomic per entry (the global version is bound by the L2's atomic throughput: 285 M RED + 285 M ATOM per 512-proof Z
// query). Pass 0 leaves the row's counts in hist_g and its slot / entry totals in row_tot; a one-block scan turns the row
// totals into row bases; pass 1 rebuilds the cursors from hist_g (padded counts for the batch-affine path) and scatters.
// The order of the entries inside a bucket differs from the global version; bucket SUMS do not depend on it.
// MEASURED ON B200 AND NOT THE DEFAULT (G16_MSM_ROWSORT=1 enables it): shared-memory atomics on random buckets are no faster in
// aggregate than the L2's (148 SMs x one returning ATOMS per ~4 cycles = 72 G/s against 95 G/s measured for the L2 path), and
// 512 blocks of 1024 threads leave a ragged second wave: sort stage 28.0 vs 13.7 ms per 1024 proofs (round 1 had found the
// same with a different kernel). Kept because it is exercised by the tests and documents the dead end.
static const int ROWSORT_T = 1024;
template <int PASS>
static __global__ void __launch_bounds__(ROWSORT_T)
msm_rowsort_kernel(MsmShape sh, const Fr* __restrict__ scalars, size_t row_stride, size_t elem_stride, const uint32_t* __restrict__ map,
                   int is_mont, uint32_t pad_mask, uint32_t* __restrict__ hist_g, uint32_t* __restrict__ row_tot,
                   const uint32_t* __restrict__ row_base, uint2* __restrict__ entries, uint32_t* __restrict__ refs, int ba_shift) {
#if defined(G16_EMU)
    uint32_t* h = reinterpret_cast<uint32_t*>(cuemu::g_dyn_smem);
#else
    extern __shared__ uint32_t h[];
#endif
    __shared__ uint32_t part[ROWSORT_T];
    __shared__ uint32_t part2[ROWSORT_T];
    const uint32_t row = blockIdx.x, t = threadIdx.x, nbk = (uint32_t)sh.nbk;
    const uint32_t ipt = (nbk + ROWSORT_T - 1) / ROWSORT_T;   // buckets per thread in the scans
    const uint32_t b0 = t * ipt < nbk ? t * ipt : nbk, b1 = b0 + ipt < nbk ? b0 + ipt : nbk;
    uint32_t* hg = hist_g + (size_t)row * nbk;
    if (PASS == 0) {
        for (uint32_t b = t; b < nbk; b += ROWSORT_T) h[b] = 0;
    } else {
        uint32_t s = 0;
        for (uint32_t b = b0; b < b1; b++) s += (hg[b] + pad_mask) & ~pad_mask;
        part[t] = s;
        __syncthreads();
        for (int off = 1; off < ROWSORT_T; off <<= 1) {   // Hillis-Steele inclusive scan of the per-thread sums
            uint32_t add = (int)t >= off ? part[t - off] : 0;
            __syncthreads();
            part[t] += add;
            __syncthreads();
        }
        uint32_t run = row_base[row] + part[t] - s;
        for (uint32_t b = b0; b < b1; b++) {
            h[b] = run;
            run += (hg[b] + pad_mask) & ~pad_mask;
        }
    }
    __syncthreads();
    const uint32_t gb = row * nbk;   // first global bucket id of the row
    for (uint32_t i = t; i < sh.n; i += ROWSORT_T) {
        const size_t sidx = (size_t)row * row_stride + (size_t)(map ? map[i] : i) * elem_stride;
        Fr raw = scalars[sidx];
        if (raw.is_zero()) continue;
        Recoded r = recode_value(raw, is_mont);
        int carry = 0;
        for (int w = 0; w < sh.nwin; w++) {
            int d = recode_digit(r, w, sh.c, carry);
            if (d == 0) continue;
            const uint32_t neg = r.neg ^ (d < 0 ? 1u : 0u);
            const uint32_t lb = (uint32_t)(d < 0 ? -d : d) - 1u;
            if (PASS == 0) {
                atomicAdd(&h[lb], 1u);
            } else {
                const uint32_t pos = atomicAdd(&h[lb], 1u);
                const uint32_t ref = (uint32_t)w * sh.n + i;
                if (!refs) {
                    entries[pos] = make_uint2(gb + lb, (ref << 1) | neg);
                } else {
                    refs[pos] = (ref << 1) | neg;
                    if ((pos & ((1u << ba_shift) - 1u)) == 0) entries[pos >> ba_shift] = make_uint2(gb + lb, (pos >> ba_shift) << 1);
                }
            }
        }
    }
    if (PASS == 0) {
        __syncthreads();
        uint32_t sp = 0, sr = 0;
        for (uint32_t b = b0; b < b1; b++) { uint32_t v = h[b]; hg[b] = v; sr += v; sp += (v + pad_mask) & ~pad_mask; }
        part[t] = sp;
        part2[t] = sr;
        __syncthreads();
        for (int off = ROWSORT_T / 2; off > 0; off >>= 1) {
            if ((int)t < off) { part[t] += part[t + off]; part2[t] += part2[t + off]; }
            __syncthreads();
        }
        if (t == 0) { row_tot[2 * row] = part[0]; row_tot[2 * row + 1] = part2[0]; }
    }
}
// one block: row_base = exclusive scan of the rows' slot totals; total[0] = slots, total[1] = entries, total[2] = slots >> ba_shift
static __global__ void __launch_bounds__(ROWSORT_T)
msm_rowscan_kernel(const uint32_t* __restrict__ row_tot, uint32_t rows, uint32_t* __restrict__ row_base, uint32_t* __restrict__ total,
                   int ba_shift) {
    __shared__ uint32_t sm[ROWSORT_T];
    __shared__ uint32_t sr[ROWSORT_T];
    __shared__ uint32_t carry;
    const uint32_t t = threadIdx.x;
    if (t == 0) carry = 0;
    uint32_t raw = 0;
    __syncthreads();
    for (uint32_t base = 0; base < rows; base += ROWSORT_T) {
        const uint32_t idx = base + t;
        const uint32_t v = idx < rows ? row_tot[2 * idx] : 0;
        raw += idx < rows ? row_tot[2 * idx + 1] : 0;
        sm[t] = v;
        __syncthreads();
        for (int off = 1; off < ROWSORT_T; off <<= 1) {
            uint32_t add = (int)t >= off ? sm[t - off] : 0;
            __syncthreads();
            sm[t] += add;
            __syncthreads();
        }
        const uint32_t c0 = carry, incl = sm[t];
        if (idx < rows) row_base[idx] = c0 + incl - v;
        __syncthreads();
        if (t == ROWSORT_T - 1) carry = c0 + incl;
        __syncthreads();
    }
    sr[t] = raw;
    __syncthreads();
    for (int off = ROWSORT_T / 2; off > 0; off >>= 1) {
        if ((int)t < off) sr[t] += sr[t + off];
        __syncthreads();
    }
    if (t == 0) { total[0] = carry; total[1] = sr[0]; total[2] = carry >> ba_shift; }
}

// ------------------------------------------------------------------------------------------------ exclusive scan (u32)
// three-kernel scan: tile sums -> scan of tile sums (single block) -> rescan tiles. Tile = 256 threads x 8 items.
#define SCAN_T 256
#define SCAN_I 8
// Padded run length and direct leftovers of a bucket with v entries on the batch-affine path (pad_mask = 2^K - 1): the run holds
// the multiple of 2^K below v, plus one more full group when the remainder exceeds pad_t; a remainder of at most pad_t entries
// skips the levels instead and is added in the XYZZ accumulation directly. A group of 2^K slots that holds one or two entries
// costs as much as a full one in the levels (0.80 ns per group of 8 against 0.17 ns per direct entry, measured per sub-batch of
// the Z query), so pad_t = 5 for K = 3 and 2 for K = 2. pad_t = 0 is the plain padding (per-row sort, small-case tests).
FD uint32_t ba_padded(uint32_t v, uint32_t pad_mask, uint32_t pad_t) {
    const uint32_t m = v & pad_mask;
    return (v & ~pad_mask) + (m > pad_t ? pad_mask + 1u : 0u);
}
FD uint32_t ba_leftover(uint32_t v, uint32_t pad_mask, uint32_t pad_t) {
    const uint32_t m = v & pad_mask;
    return m <= pad_t ? m : 0u;
}
// pad_mask = 2^K - 1 rounds every count up to a multiple of 2^K before it is summed (batch-affine path: padded runs);
// raw_sums (optional) receives the unpadded tile sums, i.e. the number of real entries.
static __global__ void scan_tile_sums(const uint32_t* __restrict__ in, size_t n, uint32_t* __restrict__ tile_sums, uint32_t pad_mask = 0,
                                      uint32_t* __restrict__ raw_sums = nullptr, uint32_t pad_t = 0, uint32_t* __restrict__ left_sums = nullptr) {
    __shared__ uint32_t sm[SCAN_T];
    __shared__ uint32_t sr[SCAN_T];
    __shared__ uint32_t sl[SCAN_T];
    size_t base = (size_t)blockIdx.x * SCAN_T * SCAN_I;
    uint32_t s = 0, r = 0, l = 0;
    for (int k = 0; k < SCAN_I; k++) {
        size_t idx = base + (size_t)k * SCAN_T + threadIdx.x;
        if (idx < n) { uint32_t v = in[idx]; r += v; s += ba_padded(v, pad_mask, pad_t); l += ba_leftover(v, pad_mask, pad_t); }
    }
    sm[threadIdx.x] = s;
    sr[threadIdx.x] = r;
    sl[threadIdx.x] = l;
    __syncthreads();
    for (int off = SCAN_T / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) {
            sm[threadIdx.x] += sm[threadIdx.x + off]; sr[threadIdx.x] += sr[threadIdx.x + off]; sl[threadIdx.x] += sl[threadIdx.x + off];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        tile_sums[blockIdx.x] = sm[0];
        if (raw_sums) raw_sums[blockIdx.x] = sr[0];
        if (left_sums) left_sums[blockIdx.x] = sl[0];
    }
}
// single block: in-place exclusive scan of `m` tile sums; writes total[0] = the grand total (sorted slots),
// total[1] = the sum of raw_sums (real entries; = total[0] without padding), total[2] = total[0] >> ba_shift
static __global__ void scan_of_sums(uint32_t* __restrict__ tile_sums, size_t m, uint32_t* __restrict__ total,
                                    const uint32_t* __restrict__ raw_sums = nullptr, int ba_shift = 0, uint32_t* __restrict__ left_sums = nullptr) {
    __shared__ uint32_t sm[SCAN_T];
    __shared__ uint32_t carry;
    __shared__ uint32_t raw[SCAN_T];
    __shared__ uint32_t left_total;
    if (threadIdx.x == 0) left_total = 0;
    if (left_sums) {   // exclusive scan of the leftover tile sums (same loop as below, kept apart for clarity: m is small)
        __shared__ uint32_t lcarry;
        if (threadIdx.x == 0) lcarry = 0;
        __syncthreads();
        for (size_t base = 0; base < m; base += SCAN_T) {
            size_t idx = base + threadIdx.x;
            uint32_t v = idx < m ? left_sums[idx] : 0;
            sm[threadIdx.x] = v;
            __syncthreads();
            for (int off = 1; off < SCAN_T; off <<= 1) {
                uint32_t add = (int)threadIdx.x >= off ? sm[threadIdx.x - off] : 0;
                __syncthreads();
                sm[threadIdx.x] += add;
                __syncthreads();
            }
            uint32_t incl = sm[threadIdx.x];
            uint32_t c0 = lcarry;
            if (idx < m) left_sums[idx] = c0 + incl - v;
            __syncthreads();
            if (threadIdx.x == SCAN_T - 1) lcarry = c0 + incl;
            __syncthreads();
        }
        if (threadIdx.x == 0) left_total = lcarry;
        __syncthreads();
    }
    {
        uint32_t r = 0;
        if (raw_sums) for (size_t i = threadIdx.x; i < m; i += SCAN_T) r += raw_sums[i];
        raw[threadIdx.x] = r;
    }
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (size_t base = 0; base < m; base += SCAN_T) {
        size_t idx = base + threadIdx.x;
        uint32_t v = idx < m ? tile_sums[idx] : 0;
        sm[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < SCAN_T; off <<= 1) {   // Hillis-Steele inclusive
            uint32_t add = (int)threadIdx.x >= off ? sm[threadIdx.x - off] : 0;
            __syncthreads();
            sm[threadIdx.x] += add;
            __syncthreads();
        }
        uint32_t incl = sm[threadIdx.x];
        uint32_t c0 = carry;
        if (idx < m) tile_sums[idx] = c0 + incl - v;
        __syncthreads();
        if (threadIdx.x == SCAN_T - 1) carry = c0 + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        uint32_t r = 0;
        for (int i = 0; i < SCAN_T; i++) r += raw[i];
        total[0] = carry;
        total[1] = raw_sums ? r : carry;
        total[2] = (carry >> ba_shift) + left_total;   // what the XYZZ accumulation walks: one entry per group + the direct leftovers
    }
}
// each tile: exclusive scan of its SCAN_T*SCAN_I items (thread-contiguous layout) + tile offset ; out may alias in
// With cursor64 (batch-affine path): out[b] receives the exclusive scan of the LEFTOVER counts instead, and cursor64[b] the
// placement word of the bucket (slot offset << 32 | leftovers << 29 | 0 entries placed), see msm_ba_place.
static __global__ void scan_tiles(const uint32_t* __restrict__ in, size_t n, const uint32_t* __restrict__ tile_offs,
                           uint32_t* __restrict__ out, uint32_t pad_mask = 0, uint32_t pad_t = 0,
                           const uint32_t* __restrict__ left_offs = nullptr, unsigned long long* __restrict__ cursor64 = nullptr) {
    __shared__ uint32_t sm[SCAN_T];
    __shared__ uint32_t sl[SCAN_T];
    size_t base = (size_t)blockIdx.x * SCAN_T * SCAN_I + (size_t)threadIdx.x * SCAN_I;
    uint32_t v[SCAN_I], lv[SCAN_I];
    uint32_t s = 0, l = 0;
    for (int k = 0; k < SCAN_I; k++) {
        const uint32_t c = (base + k < n) ? in[base + k] : 0;
        v[k] = ba_padded(c, pad_mask, pad_t);
        lv[k] = ba_leftover(c, pad_mask, pad_t);
        s += v[k];
        l += lv[k];
    }
    sm[threadIdx.x] = s;
    sl[threadIdx.x] = l;
    __syncthreads();
    for (int off = 1; off < SCAN_T; off <<= 1) {
        uint32_t add = (int)threadIdx.x >= off ? sm[threadIdx.x - off] : 0;
        uint32_t addl = (int)threadIdx.x >= off ? sl[threadIdx.x - off] : 0;
        __syncthreads();
        sm[threadIdx.x] += add;
        sl[threadIdx.x] += addl;
        __syncthreads();
    }
    uint32_t run = tile_offs[blockIdx.x] + sm[threadIdx.x] - s;
    uint32_t lrun = (left_offs ? left_offs[blockIdx.x] : 0u) + sl[threadIdx.x] - l;
    for (int k = 0; k < SCAN_I; k++) {
        if (base + k < n) {
            if (cursor64) {
                cursor64[base + k] = ((unsigned long long)run << 32) | ((unsigned long long)lv[k] << 29);
                out[base + k] = lrun;
            } else {
                out[base + k] = run;
            }
        }
        run += v[k];
        lrun += lv[k];
    }
}
// NOTE scan_tile_sums reads items tile-strided while scan_tiles reads thread-contiguous: both cover the same tile
// [blockIdx*2048, +2048), so the tile sums agree.

// ------------------------------------------------------------------------------------------------ bucket accumulation
template <class C>
FD typename C::A load_point(const typename C::A* __restrict__ bases, uint32_t ref) {
#if G16_ASM
    // 128-bit loads through the read-only path
    typename C::A p;
    const uint4* src = reinterpret_cast<const uint4*>(bases + ref);
    uint4* dst = reinterpret_cast<uint4*>(&p);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(typename C::A) / 16); k++) dst[k] = __ldg(src + k);
    return p;
#else
    return bases[ref];
#endif
}

// Thread t owns sorted entries [t*L, (t+1)*L). *total_entries is read from device memory (no host sync).
template <class C>
__global__ void __launch_bounds__(128, 4)
msm_accumulate_kernel(const typename C::A* __restrict__ bases, const uint2* __restrict__ entries,
                      const uint32_t* __restrict__ total_entries, int L, typename C::X* __restrict__ bucket_sums,
                      typename C::X* __restrict__ part_val, uint32_t* __restrict__ part_key,
                      const typename C::A* __restrict__ table = nullptr) {
    typedef typename C::X X;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t M = *total_entries;
    size_t start = t * (size_t)L;
    if (start >= M) return;
    size_t end = start + L < M ? start + L : M;
    uint32_t cur = entries[start].x;
    X acc = X::inf();
    bool first = true;
    for (size_t j = start; j < end; j++) {
        uint2 e = entries[j];
        if (e.x != cur) {
            if (first) { part_val[2 * t] = acc; part_key[2 * t] = cur; first = false; }
            else bucket_sums[cur] = acc;
            acc = X::inf();
            cur = e.x;
        }
        if (e.y == MSM_INVALID) continue;   // a slot without a point (combination-table queries: an all-zero group of wires)
        // with `table` (batch-affine path): MSM_REF_TABLE marks a direct leftover, a reference into the base table
        const bool tab = table != nullptr && (e.y & MSM_REF_TABLE);
        typename C::A p = load_point<C>(tab ? table : bases, (tab ? (e.y & ~MSM_REF_TABLE) : e.y) >> 1);
        acc.madd(p, (e.y & 1u) != 0);
    }
    if (first) {
        part_val[2 * t] = acc; part_key[2 * t] = cur;
        part_key[2 * t + 1] = MSM_INVALID;
    } else {
        part_val[2 * t + 1] = acc; part_key[2 * t + 1] = cur;
    }
}

// (Measured on B200 and not adopted: for a single request, one WARP per bucket — lanes add the bucket's ~34 entries in
// parallel and meet in a shuffle tree, no chunk-edge partials, no merge passes. The merge passes it removes cost 0.25 ms, but
// 16 384 register-heavy warps run in ~14 waves of dependent additions: accumulate 0.11 -> 2.5 ms.)
// Partial sums of buckets that straddle chunk edges form a sequence with non-decreasing keys (holes = MSM_INVALID):
// level 0 = (head, tail) of every accumulate chunk. Each merge level lets a thread reduce MSM_MERGE_C consecutive
// entries: runs strictly inside its slice are complete buckets and are written out, the first and the last run go to the
// next level. The sequence shrinks by MSM_MERGE_C/2 per level, so even a bucket that holds every point of a row (the
// {0,+-1} wire vectors of the ChaCha circuit) is summed by a log-depth tree instead of one serial chain.
static const int MSM_MERGE_C = 8;
FD size_t msm_level_len(size_t M, int L, int level) {   // entries of the sequence at `level` (0 = accumulate output)
    size_t n = 2 * ((M + L - 1) / L);
    for (int l = 0; l < level; l++) n = 2 * ((n + MSM_MERGE_C - 1) / MSM_MERGE_C);
    return n;
}
template <class C>
__global__ void __launch_bounds__(128)
msm_merge_level_kernel(const uint32_t* __restrict__ total_entries, int L, int level_in,
                       const uint32_t* __restrict__ in_key, const typename C::X* __restrict__ in_val,
                       uint32_t* __restrict__ out_key, typename C::X* __restrict__ out_val,
                       typename C::X* __restrict__ bucket_sums) {
    typedef typename C::X X;
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n_in = msm_level_len(*total_entries, L, level_in);
    size_t start = t * MSM_MERGE_C;
    if (start >= n_in) return;
    size_t end = start + MSM_MERGE_C < n_in ? start + MSM_MERGE_C : n_in;
    uint32_t cur = MSM_INVALID;
    X acc = X::inf();
    bool first = true;
    for (size_t j = start; j < end; j++) {
        uint32_t k = in_key[j];
        if (k == MSM_INVALID) continue;
        if (cur == MSM_INVALID) {
            cur = k;
            acc = in_val[j];
            continue;
        }
        if (k != cur) {
            if (first) { out_val[2 * t] = acc; out_key[2 * t] = cur; first = false; }
            else bucket_sums[cur] = acc;
            acc = in_val[j];
            cur = k;
        } else {
            acc.add(in_val[j]);
        }
    }
    if (cur == MSM_INVALID) {
        out_key[2 * t] = MSM_INVALID;
        out_key[2 * t + 1] = MSM_INVALID;
    } else if (first) {
        out_val[2 * t] = acc; out_key[2 * t] = cur;
        out_key[2 * t + 1] = MSM_INVALID;
    } else {
        out_val[2 * t + 1] = acc; out_key[2 * t + 1] = cur;
    }
}
// last level (short): the first valid entry of every key run sums the run and writes the bucket
template <class C>
__global__ void __launch_bounds__(128)
msm_merge_final_kernel(const uint32_t* __restrict__ total_entries, int L, int level_in,
                       const uint32_t* __restrict__ in_key, const typename C::X* __restrict__ in_val,
                       typename C::X* __restrict__ bucket_sums) {
    typedef typename C::X X;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n = msm_level_len(*total_entries, L, level_in);
    if (i >= n) return;
    uint32_t key = in_key[i];
    if (key == MSM_INVALID) return;
    for (size_t p = i; p > 0; p--) {
        uint32_t pk = in_key[p - 1];
        if (pk == MSM_INVALID) continue;
        if (pk == key) return;   // not the leader of this run
        break;
    }
    X acc = in_val[i];
    for (size_t j = i + 1; j < n; j++) {
        uint32_t k = in_key[j];
        if (k == MSM_INVALID) continue;
        if (k != key) break;
        acc.add(in_val[j]);
    }
    bucket_sums[key] = acc;
}

// (Measured on B200 and not adopted: capping this kernel at 168 / 128 registers for 3 / 4 resident CTAs per SM leaves the
// reduction stage at 19.0 / 19.7 ms vs 18.2 ms per 1024 proofs; a one-CTA-per-row counting sort with shared-memory
// counters instead of L2 atomics doubles the sort stage, 25.6 vs 12.7 ms.)
// ------------------------------------------------------------------------------------------------ bucket reduction tree
// Node = (R, V): R = sum of the B_k below it, V = sum (k - base)*B_k. Leaves are the buckets (weight 1..g inside a
// level-1 node), upper levels use 0-based child weights:  V = sum_i V_i + span_child * sum_i i*R_i.
// Segment = one (row, window) bucket set of `n_in` items; thread = one output node.
template <class C, int LEVEL1, int LOG_G>
__global__ void __launch_bounds__(128)
msm_tree_kernel(const typename C::X* __restrict__ in_R, const typename C::X* __restrict__ in_V, uint32_t n_in,
                uint32_t n_out, uint32_t segs, int log_span_child, typename C::X* __restrict__ out_R,
                typename C::X* __restrict__ out_V) {
    typedef typename C::X X;
    const uint32_t G = 1u << LOG_G;
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)segs * n_out) return;
    uint32_t seg = (uint32_t)(gid / n_out), node = (uint32_t)(gid % n_out);
    const X* R = in_R + (size_t)seg * n_in;
    uint32_t lo = node * G;
    uint32_t hi = lo + G < n_in ? lo + G : n_in;   // children [lo, hi)
    X run = X::inf(), tot = X::inf();
    if (LEVEL1) {
        // weights 1..g: tot accumulates run after every add
        for (uint32_t k = hi; k > lo; k--) {
            run.add(R[k - 1]);
            tot.add(run);
        }
    } else {
        const X* V = in_V + (size_t)seg * n_in;
        // weights 0..g-1: child lo (weight 0) joins run after the last tot update
        for (uint32_t k = hi; k > lo + 1; k--) {
            run.add(R[k - 1]);
            tot.add(run);
        }
        run.add(R[lo]);
        for (int d = 0; d < log_span_child; d++) tot = tot.dbl();
        for (uint32_t k = lo; k < hi; k++) tot.add(V[k]);
    }
    out_R[gid] = run;
    out_V[gid] = tot;
}
// The same node for SMALL problems (a single request, a standalone MSM of <= 2^18 points), where the time is the depth of the
// dependency chain and not the work: ONE BLOCK computes one node of up to TB children in logarithmic depth — a suffix scan of the
// children's R in shared memory (S_k = sum_{j >= k} R_j, log2 TB steps), then a tree sum of the S_k (sum_k S_k = sum_k (k+1) R_k,
// the weights 1..g of a level-1 node; without S_0 the weights 0..g-1 of an upper node) — instead of one thread walking 2 g
// dependent additions. 16384 buckets: 2 launches, ~45 dependent additions, against 6 launches and ~70 for arity 4.
template <class C, int LEVEL1, int TB>
__global__ void __launch_bounds__(TB)
msm_tree_block_kernel(const typename C::X* __restrict__ in_R, const typename C::X* __restrict__ in_V, uint32_t n_in, uint32_t n_out,
                      int log_span_child, typename C::X* __restrict__ out_R, typename C::X* __restrict__ out_V) {
    typedef typename C::X X;
#if defined(G16_EMU)
    X* sm = reinterpret_cast<X*>(cuemu::g_dyn_smem);
#else
    extern __shared__ uint4 sm_raw[];
    X* sm = reinterpret_cast<X*>(sm_raw);
#endif
    const uint32_t t = threadIdx.x;
    const uint32_t seg = blockIdx.x / n_out, node = blockIdx.x % n_out;
    const uint32_t lo = node * TB;
    const uint32_t g = lo + TB < n_in ? TB : n_in - lo;   // children [lo, lo + g)
    const X* R = in_R + (size_t)seg * n_in + lo;
    X v = t < g ? R[t] : X::inf();
    sm[t] = v;
    __syncthreads();
    for (uint32_t d = 1; d < TB; d <<= 1) {   // suffix scan
        X o = t + d < TB ? sm[t + d] : X::inf();
        __syncthreads();
        v.add(o);
        sm[t] = v;
        __syncthreads();
    }
    const X run = sm[0];                      // sum of all children
    __syncthreads();
    if (!LEVEL1 && t == 0) sm[0] = X::inf();  // upper levels: child k has weight k, so S_0 does not count
    // upper levels also sum the children's own weighted sums: that tree runs in the second half of the shared memory, on the
    // upper half of the block, at the same time as the tree over the suffix sums (different warps, different partitions)
    X* smV = sm + TB;
    if (!LEVEL1) {
        const X* V = in_V + (size_t)seg * n_in + lo;
        smV[t] = t < g ? V[t] : X::inf();
    }
    __syncthreads();
    for (uint32_t s2 = TB / 2; s2 >= 1; s2 >>= 1) {   // tree sums
        const bool doS = t < s2, doV = !LEVEL1 && t >= TB / 2 && t < TB / 2 + s2;
        const uint32_t u = t - TB / 2;
        X a;
        if (doS) { a = sm[t]; a.add(sm[t + s2]); }
        else if (doV) { a = smV[u]; a.add(smV[u + s2]); }
        __syncthreads();
        if (doS) sm[t] = a;
        else if (doV) smV[u] = a;
        __syncthreads();
    }
    X tot = sm[0];
    if (!LEVEL1 && t == 0) {
        for (int d = 0; d < log_span_child; d++) tot = tot.dbl();
        tot.add(smV[0]);
    }
    if (t == 0) {
        out_R[blockIdx.x] = run;
        out_V[blockIdx.x] = tot;
    }
}

template <class C, int LOG_G>
static void msm_tree_launch(bool level1, size_t nodes, cudaStream_t stream, const typename C::X* inR, const typename C::X* inV,
                            uint32_t n_in, uint32_t n_out, uint32_t segs, int log_span, typename C::X* outR, typename C::X* outV) {
    if (level1) {
        auto k = msm_tree_kernel<C, 1, LOG_G>;
        G16_LAUNCH(k, div_up(nodes, 128), 128, 0, stream, false, inR, inV, n_in, n_out, segs, log_span, outR, outV);
    } else {
        auto k = msm_tree_kernel<C, 0, LOG_G>;
        G16_LAUNCH(k, div_up(nodes, 128), 128, 0, stream, false, inR, inV, n_in, n_out, segs, log_span, outR, outV);
    }
}

// Horner over windows: out[row] = sum_w 2^(c*w) * S[row][w]   (S = the V of the tree roots)
template <class C>
__global__ void msm_combine_kernel(const typename C::X* __restrict__ S, uint32_t rows, int nwin, int c,
                                   typename C::X* __restrict__ out) {
    typedef typename C::X X;
    uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    X acc = X::inf();
    for (int w = nwin - 1; w >= 0; w--) {
        for (int k = 0; k < c; k++) acc = acc.dbl();
        acc.add(S[(size_t)row * nwin + w]);
    }
    out[row] = acc;
}

// The same Horner pass for G1 by a team of four warps (team.cuh): block = 128 threads = one team, lane l = row 32 b + l. The
// ~254 doublings are one dependent chain whatever the size of the MSM; a thread alone needs 3.95 us per doubling and 6.3 us
// per addition, the team 2.5 and 3.9 (B200, scripts/proto/dbl_chain.cu): 2^16 points, combine 1.04 -> 0.65 ms.
static __global__ void __launch_bounds__(128)
msm_combine_team_kernel(const G1XYZZ* __restrict__ S, uint32_t rows, int nwin, int c, G1XYZZ* __restrict__ out) {
    __shared__ Fp sm[TEAM4_SM_ELEMS];
    __shared__ uint32_t flag;
    Team4 T{sm, &flag, (int)(threadIdx.x >> 5), (int)(threadIdx.x & 31), 0};
    if (threadIdx.x == 0) flag = 0u;
    __syncthreads();
    const uint32_t row = blockIdx.x * 32 + T.lane;
    const bool live = row < rows;   // dead lanes carry the point at infinity through the same barriers
    G1XYZZ acc = G1XYZZ::inf();
    for (int w = nwin - 1; w >= 0; w--) {
        if (w != nwin - 1)
            for (int k = 0; k < c; k++) acc = team_dbl(T, acc);
        const G1XYZZ s = live ? S[(size_t)row * nwin + w] : G1XYZZ::inf();
        acc = team_add(T, acc, s, false);
    }
    if (live && T.w == 0) out[row] = acc;
}
template <class C>
static void msm_combine_launch(const typename C::X* S, uint32_t rows, int nwin, int c, typename C::X* out, cudaStream_t stream) {
    if constexpr (std::is_same<C, G1>::value) {
        static const int team = [] { const char* v = getenv("G16_MSM_COMBINE_TEAM"); return v && *v ? atoi(v) : 1; }();
        if (team) {
            G16_LAUNCH(msm_combine_team_kernel, div_up(rows, 32), 128, 0, stream, true, S, rows, nwin, c, out);
            return;
        }
    }
    auto k = msm_combine_kernel<C>;
    G16_LAUNCH(k, div_up(rows, 64), 64, 0, stream, false, S, rows, nwin, c, out);
}

template <class C>
__global__ void xyzz_to_affine_kernel(const typename C::X* __restrict__ in, uint32_t n, typename C::A* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = in[i].to_affine();
}

// table[w*n + i] = 2^(c*w) * P_i  (affine). One thread per point.
template <class C>
__global__ void msm_precompute_kernel(const typename C::A* __restrict__ pts, uint32_t n, int nwin, int c,
                                      typename C::A* __restrict__ table) {
    typedef typename C::X X;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    typename C::A p = pts[i];
    table[i] = p;
    X acc = X::from_affine(p);
    for (int w = 1; w < nwin; w++) {
        for (int k = 0; k < c; k++) acc = acc.dbl();
        table[(size_t)w * n + i] = acc.to_affine();
    }
}

// arity of the block-parallel tree node: 128 children (16 KB of shared memory for G1, 32 KB for G2). A node is a chain of
// ~3 log2(TB) dependent additions in which every warp of the block takes part, and a warp alone on its scheduler partition runs
// an addition in 6.3 us whatever its live lanes: with 256 threads two warps share each partition's multiplier and every step
// takes twice as long (2^16 points, fixed base, ncu: 280 + 256 us for the two block levels of 256 children).
template <class C> struct MsmTreeBlock { static const int LOG_TB = 7, TB = 128; };
template <> struct MsmTreeBlock<G2> { static const int LOG_TB = 7, TB = 128; };

// the batch-affine levels exist for G1 only (k_msm_ba.cu)
template <class C> constexpr bool msm_ba_available() { return std::is_same<C, G1>::value; }
static inline void msm_ba_run(MsmWorkspace<G1>& ws, const G1Affine* bases, size_t max_slots, int K, cudaStream_t stream) {
    G1Affine* lvl[MSM_BA_MAX_LEVELS] = {ws.ba_lvl[0].p, ws.ba_lvl[1].p, ws.ba_lvl[2].p};
    msm_ba_levels(bases, ws.ba_refs.p, ws.total.p, max_slots, K, lvl, ws.ba_scratch.p, stream);
}
static inline void msm_ba_run(MsmWorkspace<G2>&, const G2Affine*, size_t, int, cudaStream_t) {}

// merge tree over the partial sequence the accumulate kernel left in ws.part_*[0] (n_l0 = its upper bound); complete runs land
// in bucket_sums. Grids are sized for the upper bound, threads past the live length exit.
template <class C>
static void msm_merge_partials(MsmWorkspace<C>& ws, const uint32_t* acc_total, int L, size_t n_l0, typename C::X* bucket_sums,
                               cudaStream_t stream) {
    size_t n_up = n_l0;
    int level = 0, pp = 0;
    while (n_up > 64) {
        size_t slices = (n_up + MSM_MERGE_C - 1) / MSM_MERGE_C;
        auto k = msm_merge_level_kernel<C>;
        G16_LAUNCH(k, div_up(slices, 128), 128, 0, stream, false, acc_total, L, level, ws.part_key[pp].p,
                   ws.part_val[pp].p, ws.part_key[pp ^ 1].p, ws.part_val[pp ^ 1].p, bucket_sums);
        n_up = 2 * slices;
        level++;
        pp ^= 1;
        ws.launches++;
    }
    auto k = msm_merge_final_kernel<C>;
    G16_LAUNCH(k, div_up(n_up, 128), 128, 0, stream, false, acc_total, L, level, ws.part_key[pp].p, ws.part_val[pp].p, bucket_sums);
    ws.launches++;
}

static __global__ void msm_set_u32_kernel(uint32_t* p, uint32_t v) { *p = v; }

// Plain sums of point references grouped by row — the accumulate and merge stages of the pipeline without digits, sort and
// bucket tree: entries[j] = (row, ref), sorted by row; ref = (index into bases) << 1 | negate, MSM_INVALID = nothing.
// out[row] = Sum bases[ref >> 1] (XYZZ; infinity for a row without points). Used by the combination-table form of the
// wire-driven queries (msm_bitq.cuh), where every request has one "bucket".
template <class C>
void msm_sum_rows(MsmWorkspace<C>& ws, const typename C::A* bases, const uint2* entries, uint32_t n_entries, uint32_t rows,
                  typename C::X* out, cudaStream_t stream) {
    typedef typename C::X X;
    size_t l = (size_t)n_entries / (sm_count() * 512 * 4);
    const int L = l < 8 ? 8 : (l > 64 ? 64 : (int)l);
    const size_t max_chunks = ((size_t)n_entries + L - 1) / L;
    const size_t n_l0 = 2 * max_chunks, n_l1 = 2 * ((n_l0 + MSM_MERGE_C - 1) / MSM_MERGE_C);
    ws.total.ensure(8);
    ws.part_val[0].ensure(n_l0);
    ws.part_key[0].ensure(n_l0);
    ws.part_val[1].ensure(n_l1);
    ws.part_key[1].ensure(n_l1);
    G16_CUDA(cudaMemsetAsync(out, 0, (size_t)rows * sizeof(X), stream));
    G16_LAUNCH(msm_set_u32_kernel, 1, 1, 0, stream, false, ws.total.p + 3, n_entries);
    auto k = msm_accumulate_kernel<C>;
    G16_LAUNCH(k, div_up(max_chunks, 128), 128, 0, stream, false, bases, entries, (const uint32_t*)(ws.total.p + 3), L, out,
               ws.part_val[0].p, ws.part_key[0].p, (const typename C::A*)nullptr);
    ws.launches += 2;
    msm_merge_partials<C>(ws, ws.total.p + 3, L, n_l0, out, stream);
    G16_CHECK_LAUNCH();
}

// ------------------------------------------------------------------------------------------------ host driver
// Runs the whole pipeline on `stream`; result XYZZ per row is left in ws.result (device). No host synchronisation.
template <class C>
void msm_run(MsmWorkspace<C>& ws, const MsmShape& sh, const typename C::A* bases, const Fr* scalars, size_t row_stride,
             size_t elem_stride, const uint32_t* map, int is_mont, cudaStream_t stream, StageTimer* tm, int chunk_len) {
    typedef typename C::X X;
    const size_t nbuckets = (size_t)sh.rows * sh.buckets_per_row();
    const size_t max_entries = (size_t)sh.rows * sh.n * sh.nwin;
    if (nbuckets >= 0xFFFFFFF0ull || max_entries >= 0xFFFFFFF0ull || (size_t)sh.n * sh.nwin >= (1ull << 31))
        throw std::runtime_error("msm: problem too large for 32-bit keys");
    // Batch-affine path (msm_ba.cuh; G1 only): K pairwise affine levels in front of the XYZZ accumulation when the problem is
    // large enough to amortise three more launches and the buckets are long enough for the padding to 2^K slots to stay small
    // (nominal run length = entries per bucket for uniform digits: K = 3 from 24 on, K = 2 from 8 on). G16_MSM_BA=0 disables it,
    // G16_MSM_BA_K=1..3 forces K.
    int K = 0;
    if (msm_ba_available<C>()) {
        static const int ba_on = [] { const char* v = getenv("G16_MSM_BA"); return v && *v ? atoi(v) : 1; }();
        static const int ba_k = [] { const char* v = getenv("G16_MSM_BA_K"); return v && *v ? atoi(v) : 0; }();
        static const long ba_min = [] { const char* v = getenv("G16_MSM_BA_MIN"); return v && *v ? atol(v) : (1l << 21); }();
        const double run = (double)max_entries / (double)nbuckets;
        if (ba_on && !ws.no_ba && (long)max_entries >= ba_min) K = ba_k > 0 ? (ba_k > MSM_BA_MAX_LEVELS ? MSM_BA_MAX_LEVELS : ba_k) : (run >= 24.0 ? 3 : (run >= 8.0 ? 2 : 0));
    }
    const uint32_t pad_mask = (1u << K) - 1u;
    // per-row sort in shared memory: batches over fixed-base tables with <= 2^14 buckets per row (G16_MSM_ROWSORT: 0 (default)
    // never, 1 from 32 rows on, 2 whenever the shape allows — the switch the small-case tests use)
    static const int rowsort_mode = [] { const char* v = getenv("G16_MSM_ROWSORT"); return v && *v ? atoi(v) : 0; }();
    const size_t rowsort_smem = (size_t)sh.nbk * sizeof(uint32_t);
    const bool rowsort = sh.precomp && rowsort_smem <= 65536 && rowsort_mode > 0 && (rowsort_mode > 1 || sh.rows >= 32);
    // direct leftovers (ba_padded / ba_leftover): a remainder of at most pad_t entries of a bucket skips the levels.
    // G16_MSM_BA_LEFT=0 restores the plain padding, 1 (default) uses the measured thresholds; the placement word keeps 29 bits
    // for the entries of one bucket.
    static const int ba_left = [] { const char* v = getenv("G16_MSM_BA_LEFT"); return v && *v ? atoi(v) : 1; }();
    uint32_t pad_t = (K >= 2 && ba_left && !rowsort && (size_t)sh.n * sh.nwin < (1u << 29)) ? (K == 3 ? 5u : 2u) : 0u;
    if (pad_t && ba_left >= 2) pad_t = (uint32_t)ba_left < pad_mask ? (uint32_t)ba_left : pad_mask;   // G16_MSM_BA_LEFT=2..7: explicit threshold (tuning)
    const bool cursor64 = K > 0 && !rowsort && (size_t)sh.n * sh.nwin < (1u << 29);
    // sorted slots: every non-empty bucket is padded by at most 2^K - 1 null slots
    size_t max_slots = max_entries;
    if (K) {
        const size_t a = max_entries + (size_t)pad_mask * nbuckets, b = max_entries << K;
        max_slots = a < b ? a : b;
        if (max_slots >= 0xFFFFFFF0ull) throw std::runtime_error("msm: problem too large for 32-bit slot positions");
    }
    // what the XYZZ accumulation walks: entries, or one group sum per 2^K slots plus at most pad_t direct leftovers per bucket
    size_t max_left = (size_t)pad_t * nbuckets;
    if (max_left > max_entries) max_left = max_entries;
    const size_t acc_entries = (max_slots >> K) + max_left;
    int L = chunk_len;
    if (L <= 0) {
        // enough chunks to fill the machine a few times over, but not shorter than 8 / longer than 64 entries
        size_t want_threads = sm_count() * 512 * 4;
        size_t l = acc_entries / want_threads;
        static const int lmax = [] { const char* v = getenv("G16_MSM_LMAX"); return v && *v ? atoi(v) : 64; }();
        L = l < 8 ? 8 : (l > (size_t)lmax ? lmax : (int)l);
    }
    const size_t max_chunks = (acc_entries + L - 1) / L;
    const size_t ntiles = (nbuckets + SCAN_T * SCAN_I - 1) / (SCAN_T * SCAN_I);

    ws.counts.ensure(nbuckets);
    ws.tile_sums.ensure(3 * ntiles + 3 * (size_t)sh.rows);   // padded tile sums, the raw ones, the leftover ones (or the per-row totals / bases)
    if (cursor64) ws.cursor64.ensure(nbuckets);
    ws.total.ensure(8);
    ws.entries.ensure(acc_entries);
    ws.buckets.ensure(nbuckets);
    if (K) {
        ws.ba_refs.ensure(max_slots);
        for (int l = 0; l < K; l++) ws.ba_lvl[l].ensure(max_slots >> (l + 1));
        ws.ba_scratch.ensure(msm_ba_scratch_elems(max_slots));
    }
    const size_t n_l0 = 2 * max_chunks;                                              // level-0 partial sequence
    const size_t n_l1 = 2 * ((n_l0 + MSM_MERGE_C - 1) / MSM_MERGE_C);
    ws.part_val[0].ensure(n_l0);
    ws.part_key[0].ensure(n_l0);
    ws.part_val[1].ensure(n_l1);
    ws.part_key[1].ensure(n_l1);
    ws.result.ensure(sh.rows);

    if (tm) tm->mark(ST_MSM_SORT, stream);
    G16_CUDA(cudaMemsetAsync(ws.buckets.p, 0, nbuckets * sizeof(X), stream));
    const size_t nthreads = (size_t)sh.n * sh.rows;
    uint32_t* row_tot = ws.tile_sums.p;                  // 2 words per row
    uint32_t* row_base = ws.tile_sums.p + 2 * (size_t)sh.rows;
    if (rowsort) {
        ws.tile_sums.ensure(3 * (size_t)sh.rows + 2 * ntiles);
        row_tot = ws.tile_sums.p;
        row_base = ws.tile_sums.p + 2 * (size_t)sh.rows;
#if !defined(G16_EMU)
        static bool attr_done[64] = {};   // > 48 KB of dynamic shared memory needs the opt-in (once per device and kernel)
        int dev = 0;
        G16_CUDA(cudaGetDevice(&dev));
        if (!attr_done[dev & 63]) {
            G16_CUDA(cudaFuncSetAttribute(msm_rowsort_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            G16_CUDA(cudaFuncSetAttribute(msm_rowsort_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            attr_done[dev & 63] = true;
        }
#endif
        G16_LAUNCH(msm_rowsort_kernel<0>, sh.rows, ROWSORT_T, rowsort_smem, stream, true, sh, scalars, row_stride, elem_stride, map, is_mont, pad_mask,
                   ws.counts.p, row_tot, (const uint32_t*)row_base, ws.entries.p, (uint32_t*)nullptr, K);
        G16_LAUNCH(msm_rowscan_kernel, 1, ROWSORT_T, 0, stream, true, (const uint32_t*)row_tot, sh.rows, row_base, ws.total.p, K);
    } else {
        G16_CUDA(cudaMemsetAsync(ws.counts.p, 0, nbuckets * sizeof(uint32_t), stream));
        G16_LAUNCH(msm_digits_kernel, div_up(nthreads, 256), 256, 0, stream, false, sh, scalars, row_stride, elem_stride, map, is_mont, 0,
                   ws.counts.p, ws.entries.p);
        uint32_t* raw_sums = ws.tile_sums.p + ntiles;
        uint32_t* left_sums = cursor64 ? ws.tile_sums.p + 2 * ntiles : nullptr;
        G16_LAUNCH(scan_tile_sums, (unsigned)ntiles, SCAN_T, 0, stream, true, ws.counts.p, nbuckets, ws.tile_sums.p, pad_mask, raw_sums, pad_t, left_sums);
        G16_LAUNCH(scan_of_sums, 1, SCAN_T, 0, stream, true, ws.tile_sums.p, ntiles, ws.total.p, (const uint32_t*)raw_sums, K, left_sums);
        // batch-affine path: counts <- leftover offsets, cursor64 <- placement words; otherwise counts <- write cursors
        G16_LAUNCH(scan_tiles, (unsigned)ntiles, SCAN_T, 0, stream, true, ws.counts.p, nbuckets, ws.tile_sums.p, ws.counts.p, pad_mask, pad_t,
                   (const uint32_t*)left_sums, cursor64 ? ws.cursor64.p : (unsigned long long*)nullptr);
    }
    // the XYZZ accumulation (and the merge levels after it) walk ws.entries with the live length at `acc_total`
    const uint32_t* acc_total = ws.total.p + (K ? 2 : 0);
    if (!K) {
        if (rowsort)
            G16_LAUNCH(msm_rowsort_kernel<1>, sh.rows, ROWSORT_T, rowsort_smem, stream, true, sh, scalars, row_stride, elem_stride, map, is_mont, pad_mask,
                       ws.counts.p, row_tot, (const uint32_t*)row_base, ws.entries.p, (uint32_t*)nullptr, 0);
        else
            G16_LAUNCH(msm_digits_kernel, div_up(nthreads, 256), 256, 0, stream, false, sh, scalars, row_stride, elem_stride, map, is_mont, 1,
                       ws.counts.p, ws.entries.p);
        G16_CHECK_LAUNCH();
        if (tm) tm->mark(ST_MSM_ACC, stream);
        auto k = msm_accumulate_kernel<C>;
        G16_LAUNCH(k, div_up(max_chunks, 128), 128, 0, stream, false, bases, ws.entries.p, acc_total, L, ws.buckets.p,
                   ws.part_val[0].p, ws.part_key[0].p, (const typename C::A*)nullptr);
        ws.launches += 6;
    } else {
        G16_CUDA(cudaMemsetAsync(ws.ba_refs.p, 0xFF, max_slots * sizeof(uint32_t), stream));   // padding slots stay null references
        if (rowsort)
            G16_LAUNCH(msm_rowsort_kernel<1>, sh.rows, ROWSORT_T, rowsort_smem, stream, true, sh, scalars, row_stride, elem_stride, map, is_mont, pad_mask,
                       ws.counts.p, row_tot, (const uint32_t*)row_base, ws.entries.p, ws.ba_refs.p, K);
        else
            G16_LAUNCH(msm_digits_kernel, div_up(nthreads, 256), 256, 0, stream, false, sh, scalars, row_stride, elem_stride, map, is_mont, 2,
                       ws.counts.p, ws.entries.p, ws.ba_refs.p, K, cursor64 ? ws.cursor64.p : (unsigned long long*)nullptr);
        G16_CHECK_LAUNCH();
        if (tm) tm->mark(ST_MSM_ACC, stream);
        if (getenv("G16_MSM_BA_TRACE")) fprintf(stderr, "[msm] batch-affine K=%d rows=%u n=%u c=%d max_slots=%zu\n", K, sh.rows, sh.n, sh.c, max_slots);
        msm_ba_run(ws, bases, max_slots, K, stream);
        // one group sum per 2^K slots is left: the sorted-run accumulation over (bucket, group) keys finishes the buckets
        auto k = msm_accumulate_kernel<C>;
        G16_LAUNCH(k, div_up(max_chunks, 128), 128, 0, stream, false, (const typename C::A*)ws.ba_lvl[K - 1].p, ws.entries.p, acc_total, L,
                   ws.buckets.p, ws.part_val[0].p, ws.part_key[0].p, bases);
        ws.launches += 6 + 3 * K;
    }
    if (tm) tm->mark(ST_MSM_REDUCE, stream);
    msm_merge_partials<C>(ws, acc_total, L, n_l0, ws.buckets.p, stream);
    G16_CHECK_LAUNCH();
    if (ws.entry_log.n < 3 * (ws.log_n + 1)) {   // grow the log (rare; keeps old values)
        DevBuf<uint32_t> bigger((ws.log_n + 1) * 6 + 96);
        if (ws.log_n) G16_CUDA(cudaMemcpyAsync(bigger.p, ws.entry_log.p, ws.log_n * 12, cudaMemcpyDeviceToDevice, stream));
        G16_CUDA(cudaStreamSynchronize(stream));
        ws.entry_log = std::move(bigger);
    }
    G16_CUDA(cudaMemcpyAsync(ws.entry_log.p + 3 * ws.log_n, ws.total.p + 1, 4, cudaMemcpyDeviceToDevice, stream));      // entries = additions
    G16_CUDA(cudaMemcpyAsync(ws.entry_log.p + 3 * ws.log_n + 1, ws.total.p, 4, cudaMemcpyDeviceToDevice, stream));     // sorted slots
    G16_CUDA(cudaMemcpyAsync(ws.entry_log.p + 3 * ws.log_n + 2, ws.total.p + (K ? 2 : 1), 4, cudaMemcpyDeviceToDevice, stream));   // XYZZ accumulation entries
    ws.log_n++;
    ws.last_K = K;
    // reduction tree. Arity 32 minimises the work (2 + 3/32 additions per bucket) and is used whenever its first level
    // has enough nodes to fill the machine; a small problem (a single proof: 16384 buckets -> 512 nodes) is bound by the
    // serial chain inside a node instead (2 x arity additions per level), so it gets arity 4 and more, shorter levels.
    const uint32_t segs = sh.rows * sh.segs_per_row();
    uint32_t n_in = (uint32_t)sh.nbk;
    // A single MSM re-decides the arity at every level (its upper levels are small problems: 2^22 points, fixed base, reduce
    // 3.8 -> 2.7 ms); a batch keeps arity 32 below a wide first level (measured: 16.6 vs 17.4 ms per 1024 proofs).
    const bool tree_dynamic = sh.rows == 1;
    int log_g = ((size_t)segs * ((n_in + MSM_TREE_G - 1) / MSM_TREE_G) < sm_count() * 256) ? MSM_TREE_LOG_G_SMALL : MSM_TREE_LOG_G;
    // small problems: block-parallel nodes of arity MSM_TREE_TB (log-depth) instead of thread-serial nodes of arity 4
    static const int tree_block = [] { const char* v = getenv("G16_MSM_TREE_BLOCK"); return v && *v ? atoi(v) : 1; }();
    // at most one block per SM (G1: 128 threads = one warp per partition); G2 keeps the measured 96
    static const int tree_block_max_env = [] { const char* v = getenv("G16_MSM_TREE_BLOCK_MAX"); return v && *v ? atoi(v) : 0; }();
    const int tree_block_max = tree_block_max_env > 0 ? tree_block_max_env : (std::is_same<C, G1>::value ? (int)sm_count() : 96);
    const bool use_block = tree_block && log_g == MSM_TREE_LOG_G_SMALL;
    const int TBK = MsmTreeBlock<C>::TB;
    const X* inR = ws.buckets.p;
    const X* inV = nullptr;
    int level = 1, log_span = 0, pp = 0;
    while (true) {
        if (tree_dynamic)   // re-decide per level: the upper levels of a wide tree are small problems too
            log_g = ((size_t)segs * ((n_in + MSM_TREE_G - 1) / MSM_TREE_G) < sm_count() * 256) ? MSM_TREE_LOG_G_SMALL : MSM_TREE_LOG_G;
        // only while the level is a handful of blocks: a block node does ~4x the additions of a serial one (measured on B200:
        // 2^20 points, 128+ blocks: reduce 1.65 -> 2.77 ms; a single request, 64 blocks: 1.04 -> 0.64 ms)
        const bool block_level = (use_block || (tree_dynamic && tree_block)) && log_g == MSM_TREE_LOG_G_SMALL &&
                                 (size_t)segs * ((n_in + TBK - 1) / TBK) <= (size_t)tree_block_max;
        const int lg = block_level ? MsmTreeBlock<C>::LOG_TB : log_g;   // arity of THIS level (log_g stays the serial arity)
        uint32_t n_out = (n_in + (1u << lg) - 1) >> lg;
        ws.lvlR[pp].ensure((size_t)segs * n_out);
        ws.lvlV[pp].ensure((size_t)segs * n_out);
        if (block_level) {
            const size_t smem = (size_t)2 * TBK * sizeof(X);   // suffix sums | the children's weighted sums
#if !defined(G16_EMU)
            // > 48 KB of dynamic shared memory (G2: 64 KB) needs the opt-in, and the attribute belongs to the DEVICE the call is
            // made on: a process that drives several GPUs (g16_init_multi) sets it once on each
            static bool attr_done[64] = {};
            int dev = 0;
            G16_CUDA(cudaGetDevice(&dev));
            if (smem > 48 * 1024 && !attr_done[dev & 63]) {
                G16_CUDA(cudaFuncSetAttribute(msm_tree_block_kernel<C, 1, MsmTreeBlock<C>::TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                G16_CUDA(cudaFuncSetAttribute(msm_tree_block_kernel<C, 0, MsmTreeBlock<C>::TB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                attr_done[dev & 63] = true;
            }
#endif
            if (level == 1) {
                auto k = msm_tree_block_kernel<C, 1, MsmTreeBlock<C>::TB>;
                G16_LAUNCH(k, segs * n_out, TBK, smem, stream, true, inR, inV, n_in, n_out, log_span, ws.lvlR[pp].p, ws.lvlV[pp].p);
            } else {
                auto k = msm_tree_block_kernel<C, 0, MsmTreeBlock<C>::TB>;
                G16_LAUNCH(k, segs * n_out, TBK, smem, stream, true, inR, inV, n_in, n_out, log_span, ws.lvlR[pp].p, ws.lvlV[pp].p);
            }
        } else if (lg == MSM_TREE_LOG_G)
            msm_tree_launch<C, MSM_TREE_LOG_G>(level == 1, (size_t)segs * n_out, stream, inR, inV, n_in, n_out, segs, log_span,
                                               ws.lvlR[pp].p, ws.lvlV[pp].p);
        else
            msm_tree_launch<C, MSM_TREE_LOG_G_SMALL>(level == 1, (size_t)segs * n_out, stream, inR, inV, n_in, n_out, segs, log_span,
                                                     ws.lvlR[pp].p, ws.lvlV[pp].p);
        ws.launches++;
        inR = ws.lvlR[pp].p;
        inV = ws.lvlV[pp].p;
        n_in = n_out;
        log_span += lg;
        level++;
        pp ^= 1;
        if (n_out == 1) break;
    }
    // inV now holds one root V per segment
    if (sh.precomp) {
        G16_CUDA(cudaMemcpyAsync(ws.result.p, inV, (size_t)sh.rows * sizeof(X), cudaMemcpyDeviceToDevice, stream));
    } else {
        msm_combine_launch<C>(inV, sh.rows, sh.nwin, sh.c, ws.result.p, stream);
        ws.launches++;
    }
    G16_CHECK_LAUNCH();
    if (tm) tm->mark(-1, stream);
}

}  // namespace g16
