// Host-side decoders for the gnark v0.11.0 binary artefacts that the reference hands to InitAlgorithm
// (libraries/prover/impl/prove_impl.go:86-91 ProvingKey.ReadFrom, :102-107 NewCS(...).ReadFrom):
//   * proving key  — layout of SURVEY.md Appendix A (points stay compressed here; they are decompressed on the GPU)
//   * r1cs         — layout of SURVEY.md Appendix D (ronanh/intcomp bit-packed columns, LEB128 calldata, CBOR body)
// Pure byte shuffling: no field arithmetic happens on the host.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace g16 {

struct ParseError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct PkFile {
    uint64_t n = 0;                    // domain cardinality
    uint8_t fr_hdr[5][32];             // n^-1, w, w^-1, g, g^-1 (big-endian canonical)
    const uint8_t* g1_abd = nullptr;   // alpha, beta, delta (3 x 32 B compressed)
    const uint8_t* A = nullptr; uint32_t nA = 0;
    const uint8_t* B = nullptr; uint32_t nB = 0;
    const uint8_t* Z = nullptr; uint32_t nZ = 0;
    const uint8_t* K = nullptr; uint32_t nK = 0;
    const uint8_t* g2_bd = nullptr;    // beta2, delta2 (2 x 64 B)
    const uint8_t* B2 = nullptr; uint32_t nB2 = 0;
    uint64_t nb_wires = 0, nb_inf_a = 0, nb_inf_b = 0;
    const uint8_t* inf_a = nullptr;
    const uint8_t* inf_b = nullptr;
    uint32_t n_commit_keys = 0;
    // per BSB22 commitment: Pedersen Basis and BasisExpSigma (compressed G1, 32 B each); layout recalled from
    // gnark-crypto fr/pedersen ProvingKey.WriteTo, self-consistent with oracle/setup.py (no reference key exists)
    struct Ped { const uint8_t* basis; uint32_t n_basis; const uint8_t* basis_sigma; uint32_t n_sigma; };
    std::vector<Ped> ped;
};
PkFile parse_pk(const uint8_t* data, size_t len);

// verifying key — layout of SURVEY.md Appendix B (gnark VerifyingKey.WriteTo; libraries/verifier/impl/verify_impl.go:36-58)
struct VkFile {
    const uint8_t* alpha = nullptr;    // G1, 32 B compressed
    const uint8_t* beta1 = nullptr;    // G1
    const uint8_t* beta2 = nullptr;    // G2, 64 B compressed
    const uint8_t* gamma2 = nullptr;
    const uint8_t* delta1 = nullptr;   // G1
    const uint8_t* delta2 = nullptr;
    const uint8_t* K = nullptr; uint32_t nK = 0;
    std::vector<std::vector<uint64_t>> public_and_commitment_committed;
    struct PedVk { const uint8_t* g; const uint8_t* g_root_sigma_neg; };   // two compressed G2 points per commitment key
    std::vector<PedVk> ped;
};
VkFile parse_vk(const uint8_t* data, size_t len);

enum InstrKind : uint8_t { INS_R1C = 0, INS_HINT = 1, INS_LOOKUP = 2 };

struct CommitmentInfo {
    uint64_t commitment_index = 0;
    uint64_t nb_public_committed = 0;
    std::vector<uint32_t> private_committed;
    std::vector<uint32_t> public_and_commitment_committed;
};

struct R1csFile {
    std::vector<std::vector<uint32_t>> levels;
    std::vector<uint32_t> bp_id, cons_off, wire_off;
    std::vector<uint64_t> start;
    std::vector<uint32_t> calldata;
    std::vector<uint8_t> bp_kind;                          // per blueprint
    std::vector<std::vector<uint32_t>> bp_lookup_entries;  // per blueprint: raw EntriesCalldata (empty if not a lookup)
    std::vector<uint64_t> coeffs;                          // 4 u64 per coefficient, Montgomery
    uint64_t n_public = 0, n_secret = 0, n_internal = 0, n_constraints = 0;
    std::vector<CommitmentInfo> commitments;
    size_t n_instr() const { return bp_id.size(); }
    uint64_t n_wires() const { return n_public + n_secret + n_internal; }
};
R1csFile parse_r1cs(const uint8_t* data, size_t len);

}  // namespace g16
