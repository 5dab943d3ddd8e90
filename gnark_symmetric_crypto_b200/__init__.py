"""gnark_symmetric_crypto_b200 — B200-native Groth16/BN254 prover backend for the ChaCha20-V3 / AES-V2 circuits of
reclaimprotocol/gnark-symmetric-crypto. The compute path is the CUDA library lib/libg16b200.so (sm_100a); importing the
package does not load it, the first call does, and a missing library is an ImportError (no CPU fallback)."""
from . import _lib  # noqa: F401
from .prover import (AES_128, AES_256, CHACHA20, Groth16Context, Groth16Verifier, InitAlgorithm, InitVerifier, ProveBatch, InputParams, InputVerifyParams, MsmPlan, OutputParams,  # noqa: F401
                     Prove, ProverError, Setup, Verify, aes_witness, bsb22_challenge, decompress, field_op, g2_subgroup_check, group_op, imad_peak, msm, ntt, ntt_bench, pairing_check)
