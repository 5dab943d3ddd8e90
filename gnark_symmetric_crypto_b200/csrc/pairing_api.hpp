// Host-visible entry points of the pairing kernels (pairing.cuh, compiled in k_pairing.cu — a cold translation unit).
#pragma once
#include "common.cuh"

namespace g16 {

struct PairingConsts;
struct LineRec;

// scratch shared by successive checks on one stream (line records: 102 steps x 160 B per pair)
struct PairingWorkspace {
    DevBuf<uint8_t> consts, recs, ok;
    bool consts_ready = false;
};
// ok_out[c] (device, n_checks bytes) = 1 iff prod_{j < pairs_per_check} e(P[c*ppc + j], Q[c*ppc + j]) == 1. Affine
// Montgomery points on the device; a pair with P or Q at infinity contributes 1. Asynchronous on `st`.
// computes the constants once (ws.consts: starts with the two Frobenius constants xi^((p-1)/3), xi^((p-1)/2) as Fp2)
void pairing_consts_ensure(PairingWorkspace& ws, cudaStream_t st);
void pairing_check_run(PairingWorkspace& ws, const G1Affine* Ps, const G2Affine* Qs, uint32_t pairs_per_check, uint32_t n_checks,
                       uint8_t* ok_out, cudaStream_t st);

}  // namespace g16
