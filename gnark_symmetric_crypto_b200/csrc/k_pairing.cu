// Cold translation unit (-DG16_COLD): the pairing product check under groth16.Verify (pairing.cuh). Latency-bound chains
// of dependent field products; the Montgomery product stays out of line.
#include "pairing.cuh"
#include "pairing_api.hpp"

namespace g16 {

void pairing_consts_ensure(PairingWorkspace& ws, cudaStream_t st) {
    if (ws.consts_ready) return;
    ws.consts.ensure(sizeof(PairingConsts));
    G16_LAUNCH(pairing_consts_kernel, 1, 1, 0, st, false, (PairingConsts*)ws.consts.p);
    G16_CHECK_LAUNCH();
    ws.consts_ready = true;
}
void pairing_check_run(PairingWorkspace& ws, const G1Affine* Ps, const G2Affine* Qs, uint32_t pairs_per_check, uint32_t n_checks,
                       uint8_t* ok_out, cudaStream_t st) {
    const uint32_t n_pairs = pairs_per_check * n_checks;
    if (!n_pairs) return;
    pairing_consts_ensure(ws, st);
    ws.recs.ensure((size_t)n_pairs * PAIRING_STEPS * sizeof(LineRec));
    LineRec* recs = (LineRec*)ws.recs.p;
    G16_LAUNCH(pairing_lines_kernel, div_up(n_pairs, 64), 64, 0, st, false, Ps, Qs, n_pairs, (const PairingConsts*)ws.consts.p, recs,
               (size_t)n_pairs);
    G16_LAUNCH(pairing_check_kernel, n_checks, PAIRING_THREADS, 0, st, true, (const LineRec*)recs, (size_t)n_pairs, pairs_per_check,
               (const PairingConsts*)ws.consts.p, ok_out);
    G16_CHECK_LAUNCH();
}

}  // namespace g16
