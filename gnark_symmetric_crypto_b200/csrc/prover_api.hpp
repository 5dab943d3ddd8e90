// Host-visible types and launch wrappers of the non-MSM / non-NTT prover kernels (prover_kernels.cuh, compiled in the
// cold translation unit k_prover.cu).
#pragma once
#include "common.cuh"

namespace g16 {

static const uint32_t SOLVE_WIRE_NONE = 0xFFFFFFFFu;
static const uint32_t WIRE_CONST = 0xFFFFFFFFu;
static const uint32_t HINT_NBITS = 4115454955u, HINT_COUNT = 2138922168u, HINT_RANDOMIZE = 1774611027u,
                      HINT_BSB22 = 4156202267u;
static const int SOLVER_WARPS = 4;

struct InsMeta {
    uint32_t cd_start;     // index into calldata
    uint32_t kind;         // InstrKind | (solve_side << 8)
    uint32_t solve_wire;   // wire this R1C defines, or SOLVE_WIRE_NONE (pure check)
    uint32_t cons_off;
    uint32_t wire_off;
    uint32_t lookup_tab;   // lookup instructions: index of the 256-entry table ; countHint: offset into count_index
};

struct SolverProgram {
    const uint32_t* calldata;
    const InsMeta* meta;
    const uint32_t* level_instr;   // instruction ids grouped by level
    const uint32_t* level_off;     // nlevels + 1
    const Fr* coeffs;              // Montgomery
    const Fr* ucoef_inv;           // per instruction: 1 / (sum of the coefficients of the solved wire)
    const Fr* lookup_tabs;         // [ntab][256]
    uint32_t nlevels;
    uint32_t n_wires;
    uint32_t n_dom;                // stride of the A/B/C vectors (domain size)
    int fast_coeffs;               // coefficient ids 1,2,3,4 are +1,+2,-1,-2 (verified on the device at init)
    const Fr* randomize;           // per witness: value of the hints.Randomize wire (Montgomery), or null
    uint32_t bsb_ins;              // instruction id of the Bsb22 commitment hint (executed by the host pipeline), or ~0
    const uint32_t* count_index;   // per countHint: nq, row_pos0, row_stride_words, nrows, then nq + 1 calldata positions (or null)
};

struct AssemblyKeys {
    G1Affine alpha, beta, delta;
    G2Affine beta2, delta2;
    const G1Affine* delta_tab;    // device: j*16^i*delta,  64 x 15 entries
    const G2Affine* delta2_tab;   // device: j*16^i*delta2
};
struct AssemblyScratch {
    DevBuf<G1XYZZ> Ar, Bs1, nrsd, win_tab;   // win_tab: 15 window multiples per variable-base half-product (4 per proof)
};

// Groth16 Setup (k_setup.cu): r1cs bytes + trapdoor (6 x 32-byte big-endian: tau alpha beta gamma delta sigma) -> pk / vk bytes
void setup_run(const uint8_t* r1cs_bytes, size_t r1cs_len, const uint8_t* trapdoor_be, std::vector<uint8_t>& pk, std::vector<uint8_t>& vk,
               cudaStream_t st);

// all pointers are device pointers
void launch_decompress_g1(const uint8_t* in, uint32_t n, G1Affine* out, uint32_t* err, cudaStream_t st);
void launch_decompress_g2(const uint8_t* in, uint32_t n, G2Affine* out, uint32_t* err, cudaStream_t st);
void launch_scalars_from_be(const uint8_t* in, uint32_t n, Fr* out, cudaStream_t st);
void launch_chacha_witness(const uint8_t* keys, const uint8_t* nonces, const uint32_t* counters, const uint8_t* inputs,
                           uint32_t n, Fr* W, size_t w_stride, uint8_t* ct_out, cudaStream_t st);
void launch_witness_copy(const Fr* witness, uint32_t n_witness, uint32_t batch, Fr* W, size_t w_stride, cudaStream_t st);
// h_level_off: host copy of the level offsets (nlevels + 1). Inside level l the instructions are ordered
//   [off[l], split[l])  short expressions          -> witness-parallel kernel (a lane per witness)
//   [split[l], split2[l]) long expressions          -> term-parallel kernel (a warp per instruction and witness)
//   [split2[l], off[l+1]) logderivarg.countHint     -> block-parallel kernel (needs sp.count_index)
// (split / split2 may be null: everything is "short".) Returns the number of kernel launches. Runs levels [lev_begin, lev_end).
// cache (optional, owned by one prover context): CUDA graphs of the level launches for small batches
struct SolverGraphCache;
SolverGraphCache* solver_graph_cache_create();
void solver_graph_cache_destroy(SolverGraphCache* cache);
size_t launch_solver(const SolverProgram& sp, const uint32_t* h_level_off, const uint32_t* h_level_split,
                     const uint32_t* h_level_split2, uint32_t lev_begin, uint32_t lev_end, uint32_t batch,
                     Fr* W, size_t w_stride, Fr* A, Fr* B, Fr* C, uint32_t* status, cudaStream_t st, SolverGraphCache* cache = nullptr);
// fills the count_index entries of the given countHint instructions (device array `ids`, n of them); out = sp.count_index
void launch_solver_count_index(const SolverProgram& sp, const uint32_t* ids, uint32_t n, uint32_t* out, cudaStream_t st);
// fills ucoef_inv (n_instr entries) and returns whether coefficient ids 0..4 are 0,1,2,-1,-2 (synchronises the stream)
int launch_solver_init(const SolverProgram& sp, uint32_t n_instr, uint32_t n_coeffs, Fr* ucoef_inv, cudaStream_t st);
// builds the 64 x 15 fixed-base tables of delta / delta2 (device buffers owned by the caller)
void launch_fixed_base_tables(const AssemblyKeys& keys, G1Affine* tab1, G2Affine* tab2, cudaStream_t st);
void launch_fixed_base_table_g1(const G1Affine& base, G1Affine* tab, cudaStream_t st);   // k_assemble.cu
// G1 half of the assembly (Ar, Krs, trailer): two launches; returns the launch count. The G2 element Bs is written by
// launch_assemble_g2, which only needs the G2 MSM result and may run on another stream.
size_t launch_assemble(const AssemblyKeys& keys, AssemblyScratch& sc, bool with_commitment, uint32_t n, const G1XYZZ* mA,
                       const G1XYZZ* mB1, const G1XYZZ* mK, const G1XYZZ* mZ, const Fr* rs, uint8_t* out, size_t out_stride,
                       cudaStream_t st);
// the same in two parts (team form, G16_ASSEMBLE_TEAM != 0): what only needs the A / B1 results, and the rest
bool assemble_team_enabled();
size_t launch_assemble_products(const AssemblyKeys& keys, AssemblyScratch& sc, uint32_t n, const G1XYZZ* mA, const G1XYZZ* mB1,
                                const Fr* rs, cudaStream_t st);
size_t launch_assemble_finish(AssemblyScratch& sc, bool with_commitment, uint32_t n, const G1XYZZ* mK, const G1XYZZ* mZ, uint8_t* out,
                              size_t out_stride, cudaStream_t st);
void launch_assemble_g2(const AssemblyKeys& keys, uint32_t n, const G2XYZZ* mB2, const Fr* rs, uint8_t* out, size_t out_stride,
                        cudaStream_t st);
// AES-CTR witness (provers.go:172-227): key_len 16 or 32; W wire-major
void launch_aes_witness(const uint8_t* keys, uint32_t key_len, const uint8_t* nonces, const uint32_t* counters,
                        const uint8_t* inputs, uint32_t n, Fr* W, size_t w_stride, uint8_t* ct_out, cudaStream_t st);
// BSB22: commitment point (XYZZ per proof) -> challenge = hash_to_field(X||Y, "bsb22-commitment") written to the commitment
// wire; the affine commitment is kept for the proof
void launch_bsb22_challenge(const G1XYZZ* commit, uint32_t n, Fr* W, size_t w_stride, uint32_t commit_wire, G1Affine* commit_aff,
                            cudaStream_t st);
// proof trailer with one commitment: u32 1 | C | PoK (Appendix C) at byte 128 of each proof
void launch_assemble_commitment(const G1Affine* commit_aff, const G1XYZZ* pok, uint32_t n, uint8_t* out, size_t out_stride,
                                cudaStream_t st);
// groth16.Verify: unpack n proofs (stride bytes apart) into the pairing inputs (see prover_kernels.cuh), kSum, verdict
struct VerifyKeys {
    G1Affine alpha;
    G2Affine beta2, gamma2, delta2, ped_g, ped_gneg;
    uint32_t n_commit;
};
void launch_verify_unpack(const VerifyKeys& keys, const uint8_t* proofs, size_t stride, uint32_t n, G1Affine* P, G2Affine* Q,
                          G1Affine* P2, G2Affine* Q2, G1Affine* commit, const void* frob, uint32_t* bad, cudaStream_t st);
// frob: device pointer to the two Frobenius constants xi^((p-1)/3), xi^((p-1)/2) (the head of PairingConsts, pairing_api.hpp)
void launch_g2_subgroup(const G2Affine* pts, uint32_t n, const void* frob, uint8_t* ok, cudaStream_t st);
void launch_verify_ksum(const G1XYZZ* msm, const G1Affine* commit, uint32_t n, G1Affine* P, cudaStream_t st);
void launch_verify_verdict(const uint8_t* ok1, const uint8_t* ok2, const uint32_t* bad, uint32_t n, uint8_t* out, cudaStream_t st);
void launch_g1_affine_to_xyzz(const G1Affine* in, uint32_t n, G1XYZZ* out, cudaStream_t st);
void launch_wires_to_rows(const Fr* W, size_t w_stride, uint32_t batch, uint32_t nb_wires, Fr* out, cudaStream_t st);
// stage-level test entry points
void launch_field_op(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n, cudaStream_t st);
void launch_group_op(int group, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n, cudaStream_t st);

}  // namespace g16
