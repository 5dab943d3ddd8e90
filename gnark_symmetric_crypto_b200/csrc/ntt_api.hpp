// Host-visible types and entry points of the Fr NTT / compute_h stage (kernels: ntt.cuh, compiled in k_ntt.cu).
#pragma once
#include "common.cuh"

namespace g16 {

// Tile of one pass: 2^NTT_TILE_LG elements * 32 B of shared memory. Measured on B200 (batch of 512 x 2^15 / one 2^24):
// tile 2^11, 3 CTAs/SM 48.1 / 47.9 Gmul/s; tile 2^10, 4 CTAs/SM (64 registers) 48.8 / 49.0; 2 or 4 butterflies in flight
// per thread (93-167 registers, fewer resident warps) 41.6 / 37.0: occupancy beats per-thread ILP here.
#ifndef NTT_TILE_LG
#define NTT_TILE_LG 10
#endif
static const int NTT_MAX_TILE_LG = NTT_TILE_LG;
static const int NTT_THREADS = 256;

struct NttPass {
    int k;         // log2 n
    int b_lo;      // lowest butterfly bit of this pass
    int m;         // number of butterfly bits [b_lo, b_lo+m)
    int q;         // number of contiguous low bits carried along (strided passes); 0 when b_lo == 0
    int lg_tile;   // log2 of the tile element count
    int dif;       // 1: process bits high -> low (DIF) ; 0: low -> high (DIT)
};

struct NttDomain {
    int k = 0;
    uint32_t n = 0;
    DevBuf<Fr> tw_fwd, tw_inv;          // w^i, w^-i  (i < n/2)
    DevBuf<Fr> scale_ninv;              // 1/n
    DevBuf<Fr> scale_coset_fwd;         // g^brev(j) / n    (iNTT output -> coset coefficients)
    DevBuf<Fr> scale_coset_inv;         // g^-brev(j) / n   (coset iNTT output -> coefficients)
    DevBuf<Fr> scale_coset_only;        // g^brev(j)
    DevBuf<Fr> scale_ninv_den;          // den / n            (compute_h: coefficients of C, pre-multiplied by den = 1/(g^n - 1))
    DevBuf<Fr> scale_coset_inv_den;     // g^-brev(j) den / n (compute_h: coefficients of E)
    Fr den;                             // 1/(g^n - 1), Montgomery
    std::vector<NttPass> dif_passes, dit_passes;
    size_t launches = 0;
};

// w, g: Montgomery limbs of the n-th root of unity and of the coset generator (read from the pk header)
void ntt_domain_init(NttDomain& d, int k, const Fr& w, const Fr& g, cudaStream_t stream);
// domain for a standalone transform of size 2^k: w = root28^(2^(28-k)), g = 5 (derived on the device)
void ntt_standalone_domain(NttDomain& d, int k, cudaStream_t stream);
// DIF: natural -> bit-reversed ; DIT: bit-reversed -> natural. `scale` (optional, n entries) is indexed by the position
// on the bit-reversed side and applied there. No other normalisation.
void ntt_run(NttDomain& d, Fr* data, size_t vec_stride, uint32_t batch, bool dif, bool inverse_root, const Fr* scale,
             cudaStream_t stream);
// a, b, c: `batch` contiguous vectors of n Fr each (vec_stride == n), zero-padded evaluations in natural order. On
// return `a` holds the coefficients of H in gnark's array order (bit-reversed); b and c are clobbered.
void compute_h_run(NttDomain& d, Fr* a, Fr* b, Fr* c, size_t vec_stride, uint32_t batch, cudaStream_t stream);
// a <- d_i = A(g w^i) B(g w^i) (natural order), b clobbered: compute_h_run without its last two transforms
void compute_d_run(NttDomain& d, Fr* a, Fr* b, size_t vec_stride, uint32_t batch, cudaStream_t stream);
// rows of `out` (rows x n) <- columns j0.. of the map d -> h (which = 0) or c_evals -> -h... see k_ntt.cu
void compute_h_columns(NttDomain& d, Fr* out, uint32_t rows, uint32_t j0, int which, cudaStream_t stream);
void ntt_bitrev(const Fr* in, Fr* out, int k, cudaStream_t stream);
// big-endian canonical 32-byte scalars -> Montgomery Fr (reduced mod r)
void fr_be_to_mont(const uint8_t* d_in, uint32_t n, Fr* d_out, cudaStream_t stream);
// test / bench helpers
void ntt_fill_pattern(Fr* out, size_t total, cudaStream_t stream);
void ntt_count_mismatches(const Fr* a, const Fr* b, size_t total, uint32_t* d_count, cudaStream_t stream);
// integer-multiply microbenchmarks (ops per second)
void imad_peak_measure(double* imad_per_s, double* imad_wide_per_s, double* modmul_per_s);
// carry-chain wide MAD rate (IMAD.WIDE.U32.X) measured by the last imad_peak_measure call
double imad_chain_rate();

}  // namespace g16
