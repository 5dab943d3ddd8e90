/* g16b200 — C-ABI of the B200-native Groth16/BN254 prover backend (libg16b200.so).
 *
 * Drop-in boundary for the prove path of reclaimprotocol/gnark-symmetric-crypto. Two layers are exported:
 *
 *  (1) the OUTER ABI of the reference's own c-shared library, byte-compatible with what cgo generates from
 *      libraries/prover/libprove.go:  InitAlgorithm (:20-23), Prove (:30-47), Free (:25-28), enforce_binding (:17-18).
 *      A C / FFI caller that loads libprove.so today can load libg16b200.so instead.
 *  (2) the INNER seam that a Go shim under gnark's groth16.Prove binds with cgo:
 *      g16_init / g16_prove_witness / g16_prove_chacha_batch ..., replacing
 *      groth16.Prove(r1cs, pk, witness)                     libraries/prover/impl/provers.go:148,216
 *      ProvingKey.ReadFrom / NewCS(...).ReadFrom            libraries/prover/impl/prove_impl.go:86-91,102-107
 *      Proof.WriteTo                                        libraries/prover/impl/provers.go:152-157
 *
 * Conventions: every function returns an int status (0 = G16_OK) unless stated; nothing aborts or throws across the
 * ABI; the message of the last failure on the calling thread is available from g16_last_error(). Outputs are
 * caller-allocated unless stated. Field elements are "gnark in-memory" form: 4 x uint64 little-endian limbs in
 * Montgomery representation (R = 2^256) — exactly fr.Element / fp.Element, so a Go caller passes unsafe.Pointer(&v[0]).
 * Affine points are x|y (G1: 8 x uint64, G2: 16 x uint64 as X.A0,X.A1,Y.A0,Y.A1) = gnark's G1Affine / G2Affine;
 * (0,0) is the point at infinity. Serialized proofs are gnark's Proof.WriteTo bytes.
 * There is NO CPU fallback: every compute entry point needs a CUDA device and fails with G16_ERR_CUDA otherwise.
 */
#ifndef G16B200_H
#define G16B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G16_OK 0
#define G16_ERR_ARG 1        /* bad argument (NULL, size mismatch, unknown id) */
#define G16_ERR_PARSE 2      /* malformed pk / r1cs / json */
#define G16_ERR_CUDA 3       /* CUDA runtime error or no device */
#define G16_ERR_UNSAT 4      /* witness does not satisfy the constraint system */
#define G16_ERR_UNSUPPORTED 5
#define G16_ERR_STATE 6      /* e.g. cipher not initialised */

typedef struct g16_ctx g16_ctx;

int g16_version(void);
const char* g16_last_error(void);
int g16_device_count(int* n);

/* ------------------------------------------------------------------------------------------------------------------
 * (2) inner seam: context = one (proving key, constraint system) pair resident on one GPU.
 * Replaces prove_impl.go:86-91 (pk ReadFrom: ~89k G1 + 12.5k G2 decompressions, done on the GPU here) and :102-107.
 * ------------------------------------------------------------------------------------------------------------------ */
int g16_init(const uint8_t* pk, size_t pk_len, const uint8_t* r1cs, size_t r1cs_len, int device, g16_ctx** out);
void g16_free(g16_ctx* ctx);
/* Multi-GPU handle (SURVEY 8b `g16_init(..., devices[], n, &ctx)`, 8e): the same key resident on every device of the list,
 * one CUDA context + host thread per device. Independent proofs are the sharding unit: the batch entry points below send
 * request i to devices[i mod n], run the devices concurrently and gather proofs and ciphertexts in input order. There is no
 * inter-GPU traffic and no collective. A batch of one request runs on devices[0]. A device index may be listed more than
 * once (two contexts on one GPU). g16_prove_witness and the stage-level entry points use devices[0]. */
int g16_init_multi(const uint8_t* pk, size_t pk_len, const uint8_t* r1cs, size_t r1cs_len, const int* devices,
                   size_t n_devices, g16_ctx** out);
/* *n_out = number of devices of the handle; devices_out[0..min(cap, n)) = their CUDA indices (may be NULL) */
int g16_ctx_devices(const g16_ctx* ctx, int* devices_out, size_t cap, size_t* n_out);
/* borrowed single-device handle of slot k (0 = the handle itself) for callers that schedule the devices themselves, as the
 * Prove batcher does (one worker per GPU over a shared queue); owned by `ctx`, never pass it to g16_free */
int g16_ctx_device_handle(g16_ctx* ctx, size_t k, g16_ctx** out);

/* info[0..15]: 0 domain size n, 1 |G1.A|, 2 |G1.B|, 3 |G1.Z|, 4 |G1.K|, 5 |G2.B|, 6 nbWires, 7 nbPublic (incl. ONE),
 * 8 nbSecret, 9 nbConstraints, 10 nbInstructions, 11 nbLevels, 12 nbCommitments, 13 proof bytes, 14 device, 15 reserved */
int g16_info(const g16_ctx* ctx, uint64_t info[16]);

/* gnark-shaped entry: `witness` = the nbPublic-1 public then nbSecret secret assignments (what frontend.NewWitness
 * yields, provers.go:144), Montgomery limbs. rs = r|s as 2 x 32-byte big-endian canonical scalars (followed by the 32-byte
 * commitment mask for circuits with a hints.Randomize wire, i.e. AES: 96 bytes), or NULL to draw them from the OS CSPRNG
 * (gnark draws from crypto/rand). proof_out must hold g16_info()[13] bytes. */
int g16_prove_witness(g16_ctx* ctx, const uint64_t* witness, size_t n_witness, const uint8_t* rs, uint8_t* proof_out,
                      size_t* proof_len);

/* library-shaped batch entry for cipher "chacha20" (provers.go:79-158 for n independent requests): computes the
 * ciphertext, the witness, and the proof of each request on the GPU.
 * keys n*32, nonces n*12, counters n (host order), inputs n*64, rs n*64 (or NULL) -> proofs n*164, ciphertexts n*64. */
int g16_prove_chacha_batch(g16_ctx* ctx, size_t n, const uint8_t* keys, const uint8_t* nonces, const uint32_t* counters,
                           const uint8_t* inputs, const uint8_t* rs, uint8_t* proofs_out, uint8_t* ct_out);

/* Requests that the circuit rejects (an unsatisfiable witness, e.g. an AES counter whose 4-block range wraps) do not stop
 * the batch: every other proof is produced, the rejected slots are zero-filled, the call returns G16_ERR_UNSAT and
 * status_out[i] (bit 0: unsatisfied constraint, bit 1: division by zero in the solver) tells which requests failed. */
int g16_last_batch_status(g16_ctx* ctx, uint32_t* status_out, size_t n);

/* The same batch split into its three phases so a benchmark can time the device part with inputs already resident in
 * HBM: stage (H2D) -> run (all kernels, returns after the stream drained; device time in *ms if non-NULL) -> fetch (D2H). */
int g16_chacha_batch_stage(g16_ctx* ctx, size_t n, const uint8_t* keys, const uint8_t* nonces, const uint32_t* counters,
                           const uint8_t* inputs, const uint8_t* rs);
int g16_chacha_batch_run(g16_ctx* ctx, float* ms);
int g16_chacha_batch_fetch(g16_ctx* ctx, uint8_t* proofs_out, uint8_t* ct_out);

/* library-shaped batch entry for "aes-128-ctr" / "aes-256-ctr" (provers.go:172-227): key_len 16 or 32. rsm = n x 96 bytes
 * (r | s | commitment mask, 32-byte big-endian each) or NULL for CSPRNG values. proofs n*196 (one BSB22 commitment),
 * ciphertexts n*64. g16_aes_batch_stage + g16_chacha_batch_run + g16_chacha_batch_fetch split it into phases. */
int g16_prove_aes_batch(g16_ctx* ctx, size_t n, const uint8_t* keys, size_t key_len, const uint8_t* nonces,
                        const uint32_t* counters, const uint8_t* inputs, const uint8_t* rsm, uint8_t* proofs_out,
                        uint8_t* ct_out);
int g16_aes_batch_stage(g16_ctx* ctx, size_t n, const uint8_t* keys, size_t key_len, const uint8_t* nonces,
                        const uint32_t* counters, const uint8_t* inputs, const uint8_t* rsm);

/* Batch schedule. pipeline = 0 (default): one main stream with per-stage timers (g16_last_stage_ms), sub-batches of
 * `sub_batch` proofs (default 512; env G16_SUBBATCH) through the H / Z stages, wire-driven MSMs on a side stream.
 * pipeline = 1 (env G16_PIPELINE=1): sub-batches alternate between two CUDA streams so that the latency-bound phases of
 * one sub-batch overlap the multiply-bound phases of the other; only the total time is measured. On B200 this is not
 * faster (the hot kernels fill every SM's register file), so it is off by default. sub_batch = 0 keeps
 * the current value. Results are identical under every schedule.
 * Z query (circuits without commitment): batches of >= 128 proofs (env G16_EVAL_Z_MIN; G16_EVAL_Z=0/1 forces the choice)
 * never materialise H: four transforms, then MultiExp over the coset products and the C evaluations against
 * evaluation-basis tables derived once per context from pk.G1.Z. Same group element as gnark's
 * MultiExp(pk.G1.Z, h) (prove.go:267-275), same proof bytes; g16_prove_witness_detail with h_out keeps the coefficient path.
 * The side stream runs at the highest stream priority (env G16_SIDE_PRIO=0 for default priority). */
int g16_set_schedule(g16_ctx* ctx, int pipeline, int sub_batch);

/* per-stage timings of the last g16_chacha_batch_run, milliseconds: 0 witness+solve, 1 compute_h (transforms + pointwise; also hosts side-stream kernels),
 * 2 MSM scalar prep + sort, 3 MSM bucket accumulation, 4 MSM reductions, 5 proof assembly, 6 total, 7 kernel launches */
int g16_last_stage_ms(const g16_ctx* ctx, float ms[8]);
/* work counters of the last run: 0 G1 mixed additions done by the bucket-accumulation kernel (= sorted entries),
 * 1 G2 mixed additions, 2 G1 accumulate launches, 3 G2 accumulate launches, 4 kernel launches, 5 proofs,
 * 6 the part of [0] that belongs to the Z query (what the stage timers of the pipeline = 0 schedule bracket),
 * 7 sub-batch size | (1 << 32 if the pipelined schedule ran) | (1 << 33 if the Z query ran over the evaluation basis) */
int g16_last_counters(const g16_ctx* ctx, uint64_t out[8]);
/* the same eight, then (device 0 of a multi-device handle): 8 sorted slots of the Z query = [6] plus the null slots that pad
 * every bucket's run to a multiple of 2^K for the batch-affine levels, 9 K = number of batch-affine levels the Z query ran
 * (0: XYZZ accumulation only), 10 bucket sets x buckets of the Z query, 11..15 reserved */
int g16_last_counters_ex(const g16_ctx* ctx, uint64_t out[16]);

/* ------------------------------------------------------------------------------------------------------------------
 * stage-level entry points (parity tests against the oracle; standalone MSM / NTT sweeps of BASELINE config 5)
 * ------------------------------------------------------------------------------------------------------------------ */
/* field: 0 = Fp, 1 = Fr. op: 0 add, 1 sub, 2 mul, 3 inv(a), 4 sqr(a), 5 neg(a), 6 to_mont(a), 7 from_mont(a) */
int g16_field_op(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);
/* group: 1 = G1, 2 = G2. op 0: out = a + b (affine in/out) ; op 1: out = k * a, k = canonical scalar limbs in b (4 u64)
 * ; op 2: out = 2a ; op 3: a + b through the general (projective) addition ; G1 only, the four-warp forms the serial tails of
 * the MSM and of the proof assembly use (csrc/team.cuh): op 4: a + b ; op 5: 4a + 2b */
int g16_group_op(int group, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);
/* decompression of gnark-crypto compressed points (32 / 64 bytes each) -> affine Montgomery */
int g16_decompress(int group, const uint8_t* in, uint64_t* out, size_t n);

/* ---- Setup. Replaces keygen.go:359-435 (groth16.Setup(r1cs) + ProvingKey.WriteTo / VerifyingKey.WriteTo)
 * for a constraint system in gnark's serialization (r1cs.chacha20 / r1cs.aes128 / r1cs.aes256): the QAP is evaluated at the
 * trapdoor on the host, every group element of the two keys is a fixed-base product on the GPU, the keys come back in the
 * layouts InitAlgorithm / InitVerifier read. trapdoor_be = tau | alpha | beta | gamma | delta | sigma (6 x 32-byte big-endian,
 * reduced, non-zero; sigma is the Pedersen trapdoor of the BSB22 commitment key) or NULL to draw it from the OS CSPRNG as gnark
 * does. *pk_out / *vk_out are malloc'd: release them with Free. */
int g16_setup(const uint8_t* r1cs, size_t r1cs_len, const uint8_t* trapdoor_be, int device, uint8_t** pk_out, size_t* pk_len,
              uint8_t** vk_out, size_t* vk_len);

/* ---- verifier: groth16.Verify for batches of proofs of one verifying key.
 * Replaces gnark v0.11.0 backend/groth16/bn254/verify.go Verify + VerifyingKey.ReadFrom as used by
 * libraries/verifier/impl/verify_impl.go:36-58 and verifiers.go:87-99,133-145.
 * vk: the bytes of VerifyingKey.WriteTo (vk.chacha20 / vk.aes128 / vk.aes256). proofs: n x proof_bytes (Proof.WriteTo, 164 B
 * without / 196 B with one BSB22 commitment). public_inputs: n x n_public field elements (the public witness without ONE),
 * public_format 0 = gnark in-memory fr.Element (4 x u64 LE Montgomery), 1 = 32-byte big-endian canonical.
 * ok_out[i] = 1 iff proof i is accepted; malformed proofs are rejected (0), they do not fail the call. */
typedef struct g16_vctx g16_vctx;
int g16_verify_init(const uint8_t* vk, size_t vk_len, int device, g16_vctx** out);
int g16_verify_info(const g16_vctx* ctx, uint64_t info[4]);   /* n_public, n_commitments, proof_bytes, len(G1.K) */
int g16_verify_batch(g16_vctx* ctx, size_t n, const uint8_t* proofs, const void* public_inputs, int public_format,
                     uint8_t* ok_out, float* device_ms);
void g16_verify_free(g16_vctx* ctx);

/* Pairing product check, the arithmetic under groth16.Verify (gnark-crypto v0.14.0 ecc/bn254 PairingCheck, called by gnark
 * v0.11.0 backend/groth16/bn254/verify.go from libraries/verifier/impl/verifiers.go:93-99,139-145).
 * ok_out[c] = 1 iff prod_{j < pairs_per_check} e(P[c*ppc + j], Q[c*ppc + j]) == 1. Points: affine Montgomery, G1 8 x u64,
 * G2 16 x u64 (x.a0 x.a1 y.a0 y.a1), (0,0) = infinity (such a pair contributes 1). */
int g16_pairing_check(const uint64_t* g1_points, const uint64_t* g2_points, size_t pairs_per_check, size_t n_checks,
                      uint8_t* ok_out);
/* ok_out[i] = 1 iff the affine twist point i lies in the r-torsion subgroup G2 (gnark-crypto ecc/bn254/g2.go IsInSubGroup, the
 * check G2Affine.SetBytes applies to every decoded point; g16_verify_batch applies it to Bs). */
int g16_g2_subgroup_check(const uint64_t* g2_points, size_t n, uint8_t* ok_out);

/* Pippenger MSM over n affine points (replaces (*G1Jac).MultiExp / (*G2Jac).MultiExp).
 * scalars: n x 4 u64; scalars_mont != 0 if they are in Montgomery form (gnark passes fr.Element vectors).
 * window = 0 picks c automatically. out = affine result. timing (optional): ms[0] total device time,
 * ms[1] bucket-accumulation kernel, ms[2] sort (digits+scan+scatter), ms[3] reductions. Host buffers. */
int g16_msm(int group, const uint64_t* points, const uint64_t* scalars, int scalars_mont, size_t n, int window,
            uint64_t* out, float ms[4]);
/* device-resident variant for benchmarking: upload once, run many times */
/* Stage-level view of the combination-table form of the wire-driven queries (csrc/k_bitq.cu; gnark prove.go:197-260 MultiExp over
 * pk.G1.A / G1.B / G1.K / G2.B for wires that only hold 0, 1 or -1): out[row] = Sum_i wires[i][row] * points[i], wires wire-major
 * [n][rows] Montgomery Fr; the first n_binary wires hold 0 / 1 (groups of 8), the others 0 / 1 / -1 (groups of 5); any other
 * value sets *exception_out = 1 and the sums are meaningless. Parity tests only. */
int g16_bitq_sum(int group, const uint64_t* points, size_t n, size_t n_binary, const uint64_t* wires, size_t rows, uint64_t* out,
                 uint32_t* exception_out);

typedef struct g16_msm_plan g16_msm_plan;
int g16_msm_plan_create(int group, const uint64_t* points, size_t n, int window, int device, g16_msm_plan** out);
/* bases fixed across many MSMs (an SRS): tabulate 2^(c w) P_i once (window 0 = choose). Results are unchanged. */
int g16_msm_plan_precompute(g16_msm_plan* plan, int window);
int g16_msm_plan_set_scalars(g16_msm_plan* plan, const uint64_t* scalars, int scalars_mont);
int g16_msm_plan_run(g16_msm_plan* plan, uint64_t* out, float ms[4]);
void g16_msm_plan_free(g16_msm_plan* plan);

/* Fr NTT of size n = 2^k on host data (replaces fft.(*Domain).FFT / FFTInverse). Natural order in and
 * out; inverse != 0 applies w^-1 and the 1/n scaling; coset != 0 evaluates on / interpolates from the coset 5*<w>.
 * ms (optional) = device time of the transform kernels only. */
int g16_ntt(uint64_t* data, size_t n, int inverse, int coset, float* ms);
/* device-resident batched variant: `batch` vectors of size n, repeated `iters` times forward+inverse; returns ms/iter */
int g16_ntt_bench(size_t n, size_t batch, int iters, float* ms_per_iter, uint64_t* checksum);

/* H = (A.B - C)/Z (replaces prove.go:computeH): a,b,c = nbConstraints evaluations each.
 * h_out = n coefficients in gnark's array order (bit-reversed), which pairs index-for-index with pk.G1.Z.
 * Precondition (what the solver guarantees, and what makes the quotient a polynomial at all): c[i] == a[i] * b[i] for every
 * constraint. Under it the result is bit-identical with gnark's; it is obtained from six transforms instead of seven
 * (C never visits the coset, see csrc/k_ntt.cu compute_h_run). */
int g16_compute_h(g16_ctx* ctx, const uint64_t* a, const uint64_t* b, const uint64_t* c, uint64_t* h_out);
/* R1CS solve (replaces constraint/bn254 (*system).Solve) for `batch` independent witnesses.
 * witness: batch x n_witness. Outputs (any may be NULL): W batch x nbWires, A/B/C batch x nbConstraints. */
int g16_solve(g16_ctx* ctx, const uint64_t* witness, size_t n_witness, size_t batch, uint64_t* W, uint64_t* A,
              uint64_t* B, uint64_t* C);
/* same, for circuits with a hints.Randomize wire (AES): masks_be = batch x 32-byte big-endian commitment masks */
int g16_solve_ex(g16_ctx* ctx, const uint64_t* witness, size_t n_witness, size_t batch, const uint8_t* masks_be,
                 uint64_t* W, uint64_t* A, uint64_t* B, uint64_t* C);
/* AES-CTR witness assignment alone (provers.go:172-227: crypto/aes + cipher.NewCTR for the ciphertext, then the
 * byte-valued witness): ct_out n x 64; witness_out (may be NULL) n x (142 + key_len) field elements, Montgomery, in
 * solver order ONE | Nonce[12] | Counter | Plaintext[64] | Ciphertext[64] | Key[key_len]. */
int g16_aes_witness(const uint8_t* keys, size_t key_len, const uint8_t* nonces, const uint32_t* counters,
                    const uint8_t* inputs, size_t n, uint8_t* ct_out, uint64_t* witness_out);
/* BSB22 commitment challenge alone (gnark prove.go:84-108 -> gnark-crypto fr.Hash / hash_to_field, RFC 9380
 * expand_message_xmd with SHA-256, DST "bsb22-commitment"): n affine G1 commitments (Montgomery x|y) -> n Fr, Montgomery */
int g16_bsb22_challenge(const uint64_t* commitments, size_t n, uint64_t* challenges_out);
/* the five MSM results of one proof before assembly: affine msmA, msmB1, msmK, msmZ (G1) and msmB2 (G2) */
int g16_prove_witness_detail(g16_ctx* ctx, const uint64_t* witness, size_t n_witness, const uint8_t* rs,
                             uint8_t* proof_out, size_t* proof_len, uint64_t* msm_g1_out /* 4 x 8 */,
                             uint64_t* msm_g2_out /* 16 */, uint64_t* h_out /* n x 4 or NULL */);

/* integer-multiply microbenchmark: sustained 32-bit IMAD and IMAD.WIDE rates of this GPU (ops/s), the roofline
 * denominator the MSM numbers are quoted against (not in MEASURED_PEAKS.json). */
int g16_imad_peak(double* imad_per_s, double* imad_wide_per_s, double* modmul_per_s);
/* rate of carry-chained wide MADs (IMAD.WIDE.U32.X, the instruction the Montgomery product is made of) measured by the
 * last g16_imad_peak call; 0 before */
int g16_imad_chain_rate(double* wide_carry_mads_per_s);

/* ------------------------------------------------------------------------------------------------------------------
 * (1) outer ABI: byte-compatible with the cgo exports of libraries/prover/libprove.go
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct { void* data; long long len; long long cap; } GoSlice_g16;
typedef struct { void* r0; long long r1; } Prove_return_g16;
void enforce_binding(void);                                                        /* libprove.go:17-18 */
/* Devices: env G16_DEVICES = comma-separated CUDA indices or "all" (default: G16_DEVICE, else 0). With more than one device
 * the key is loaded on each and concurrent Prove calls are served by one batching worker per GPU over a shared queue. */
unsigned char InitAlgorithm(unsigned char algorithmID, GoSlice_g16 provingKey, GoSlice_g16 r1cs);   /* libprove.go:20-23 */
void Free(void* pointer);                                                          /* libprove.go:25-28 */
Prove_return_g16 Prove(GoSlice_g16 params);                                        /* libprove.go:30-47 */
/* verifier side, libraries/verifier/libverify.go:14-17: Verify(params JSON {"cipher","proof","publicSignals"}) -> bool.
 * The reference embeds vk.chacha20 / vk.aes128 / vk.aes256 with go:embed (impl/verify_impl.go:26-62); this library takes the
 * same bytes once per cipher through InitVerifier (algorithm ids as InitAlgorithm). */
/* SURVEY 8f rank 3, beside Prove (which is unchanged): params = a JSON array of InputParams objects (provers.go:53-59), the
 * result = a JSON array with one element per request, in order: the OutputParams object Prove would return
 * (prove_impl.go:49-52) or, for a request Prove would have panicked on, the JSON-encoded error string. Requests of any mix
 * of initialised ciphers; each cipher's share is proved as one batch (on every GPU of G16_DEVICES). A payload that is not
 * a JSON array comes back as one JSON string (the panic convention of libprove.go:33-43). Release with Free. */
Prove_return_g16 ProveBatch(GoSlice_g16 params);
/* counters of the Prove batcher of one cipher: out[0] batches run, out[1] requests proved, out[2] devices serving it */
int g16_libprove_stats(int algorithmID, uint64_t out[3]);
unsigned char InitVerifier(unsigned char algorithmID, GoSlice_g16 verifyingKey);
unsigned char Verify(GoSlice_g16 params);

#ifdef __cplusplus
}
#endif
#endif /* G16B200_H */
