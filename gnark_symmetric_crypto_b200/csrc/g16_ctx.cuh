// Prover context: one (proving key, constraint system) pair resident on one GPU, and the batched prove pipeline.
// Host code only — every kernel is launched through the wrappers of msm_types.hpp / ntt_api.hpp / prover_api.hpp.
// Replaces gnark v0.11.0 backend/groth16/bn254/prove.go:64-295 (Prove) and marshal.go:311-348 (ProvingKey.ReadFrom),
// as driven by libraries/prover/impl/prove_impl.go:65-114 (InitAlgorithm) and provers.go:79-158 (proveChaCha).
#pragma once
#include "common.cuh"
#include "host_parse.hpp"
#include "msm_types.hpp"
#include "ntt_api.hpp"
#include "pairing_api.hpp"
#include "prover_api.hpp"
#include <memory>

namespace g16 {

struct PrecompQuery {
    DevBuf<G1Affine> table;
    DevBuf<uint32_t> map;   // scalar index (wire id) of point i ; empty = identity
    uint32_t n = 0;
    int c = 0;
};

// Combination-table form of one wire-driven query (k_bitq.cu): the points whose wire is a bit, in groups of 8 with the table of
// the 255 subset sums of every group, and the remaining points as a general sub-query over gathered window tables.
struct BitQuery {
    bool on = false;
    uint32_t groups = 0, groups_bin = 0;   // groups [0, groups_bin): 8 wires in {0, 1}; the others: 5 wires in {0, 1, -1}
    DevBuf<uint32_t> grp_wires;     // 8 slots of wire ids per group (BITQ_NONE pads)
    DevBuf<G1Affine> table1;        // [groups][256] subset sums on G1
    DevBuf<G2Affine> table2;        // the same on G2 (B query only)
    PrecompQuery rest;              // points that are not bits: general path (map = their wires, table = gathered windows)
    DevBuf<G2Affine> rest_tab2;     // ... their G2 window tables (B query only)
};

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    // sizes
    uint64_t n_dom = 0;
    int k_dom = 0;
    uint32_t nA = 0, nB = 0, nZ = 0, nK = 0, nB2 = 0;
    uint32_t nb_wires = 0, n_public = 0, n_secret = 0, n_constraints = 0, n_instr = 0, nlevels = 0, n_commit = 0;
    // key material
    DevBuf<G1Affine> A, B, Z, K, ped_basis, ped_basis_sigma;
    DevBuf<G2Affine> B2;
    PrecompQuery qA, qB, qZ, qK;
    // evaluation-basis form of the Z query (ctx_build_eval_tables): Sum_j d_j Qd_j + Sum_j c_j Qc_j = Sum_k h_k Z_k
    PrecompQuery qQd, qQc;
    int eval_z = -1;                  // G16_EVAL_Z: 0 never, 1 always, -1 (default): for batches >= eval_z_min
    uint32_t eval_z_min = 128;         // measured on B200: slower below 64 proofs (+0.5 ms at n = 1), faster from 128 on
    bool eval_ready = false;
    bool want_h = false;              // the caller reads H back: stay on the coefficient-basis path
    DevBuf<G1XYZZ> resZc;
    DevBuf<G2Affine> tabB2;
    int cB2 = 0;
    AssemblyKeys keys;
    AssemblyScratch asm_scratch;
    DevBuf<G1Affine> delta_tab;
    DevBuf<G2Affine> delta2_tab;
    NttDomain dom;
    // solver program
    DevBuf<uint32_t> d_calldata, d_level_instr, d_level_off, d_count_index;
    DevBuf<InsMeta> d_meta;
    DevBuf<Fr> d_coeffs, d_ucoef_inv, d_lookup_tabs;
    SolverProgram sp;
    std::vector<uint32_t> h_level_off, h_level_split, h_level_split2;   // split: first "long" instruction of each level (see ctx_create)
    bool solver_supported = true;
    std::string solver_unsupported_reason;
    // BSB22 commitment (AES circuits): at most one commitment is supported
    PrecompQuery qPed, qPedSigma;       // Pedersen Basis / BasisExpSigma over the private-committed wires
    uint32_t commit_wire = 0, bsb_level = 0, bsb_ins = 0xFFFFFFFFu;
    bool has_randomize = false;
    DevBuf<G1XYZZ> resCommit, resPok;
    DevBuf<G1Affine> commit_aff;
    DevBuf<Fr> d_mask;
    DevBuf<uint8_t> d_mask_be;
    // batch state
    size_t staged = 0;
    int staged_kind = 0;   // 0 generic witness, 1 chacha requests, 2 aes requests
    uint32_t staged_key_len = 0;
    DevBuf<uint8_t> d_keys, d_nonces, d_inputs, d_rs_be, d_ct, d_proofs;
    DevBuf<uint32_t> d_counters, d_status;   // d_status: one word per witness of the batch (solver.cuh status bits)
    std::vector<uint32_t> h_status;          // the same, on the host after the last run (g16_last_batch_status)
    DevBuf<Fr> d_rs, d_witness, W, Aev, Bev, Cev;
    MsmWorkspace<G1> ws1;    // Z query, main stream (lane 0)
    MsmWorkspace<G1> ws1c;   // Z query of lane 1 (pipelined schedule)
    MsmWorkspace<G1> ws1b;   // A / B1 / K queries, side stream
    MsmWorkspace<G2> ws2;    // B2 query, side stream
    cudaStream_t stream2 = nullptr;
    cudaStream_t stream3 = nullptr;   // lane 1 of the pipelined schedule
    cudaStream_t stream4 = nullptr;   // the half-products of the assembly, as soon as the A and B1 queries of the batch are done
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join3 = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_ab = nullptr, ev_prod = nullptr;   // A / B1 results complete (side stream) ; half-products done (stream4)
    // small batches: the four wire-driven queries run side by side (ctx_wire_queries_fan) — B1, B2, K on streams of their own
    cudaStream_t fan[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_fan[4] = {nullptr, nullptr, nullptr, nullptr};
    MsmWorkspace<G1> wsf[2];   // B1 and K queries of the fan (ws1b keeps A)
    std::vector<cudaEvent_t> ev_solved;   // one per sub-batch: witness complete
    std::vector<cudaEvent_t> ev_hdone;    // pipelined schedule: transforms of sub-batch k done
    bool pipeline_stagger = true;         // G16_PIPE_STAGGER
    bool test_lane_oom = false;           // G16_TEST_LANE_OOM
    int pipeline = 0;                 // 1: sub-batches alternate between two lanes (streams); 0: one main stream, stage timers
    DevBuf<G1XYZZ> resA, resB1, resK, resZ;
    DevBuf<G2XYZZ> resB2;
    StageTimer timer;
    float stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t launches = 0;
    uint64_t counters[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t sub_batch = 512;
    size_t solve_chains = 2;           // G16_SOLVE_CHAINS: concurrent solver chains of mode 2 (2 = one per sub-batch on the two streams)
    std::vector<cudaStream_t> solve_streams;
    int split_solve = 0;               // G16_SPLIT_SOLVE: 1 = later sub-batches solved on the side stream behind the first one's transforms, 2 = all sub-batches solved concurrently (ctx_run_batch)
    bool tables_ready = false;
    SolverGraphCache* solver_graphs = nullptr;
    // combination tables of the wire-driven queries (A, B1, K, B2). Which wires are bits is learned from the first
    // bitq_min_rows witnesses (state 0), then the tables are built (state 1) and used for batches >= bitq_min_batch; every
    // witness is checked against the classification, and an exception sends the batch through the general path again and
    // switches the tables off for good (state 2). G16_BITQ=0 disables the path.
    int bitq_state = 0;
    bool bitq_built = false;
    size_t bitq_rows_seen = 0;
    uint32_t bitq_min_rows = 256, bitq_min_batch = 32;
    std::vector<uint32_t> bit_mask;                // host, per wire: bit 0 while every witness seen so far held 0 / 1, bit 1 while 0 / 1 / -1
    std::vector<uint32_t> h_mapA, h_mapB, h_mapK;
    DevBuf<uint32_t> d_bit_flags;
    DevBuf<uint32_t> d_bitq_exc;
    DevBuf<uint2> bitq_entries;
    DevBuf<G1XYZZ> bitq_tmp1;
    DevBuf<G2XYZZ> bitq_tmp2;
    BitQuery bqA, bqB, bqK;
    int bitq_relearned = 0;                        // exceptions met so far (each one re-opens the learning phase; > 3: state 2)
    bool bitq_test_exception = false;              // G16_BITQ_TEST_EXC=1: pretend one exception (tests of the fallback)

    ~Ctx() {
        solver_graph_cache_destroy(solver_graphs);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (ev_join3) cudaEventDestroy(ev_join3);
        if (ev_t0) cudaEventDestroy(ev_t0);
        if (ev_t1) cudaEventDestroy(ev_t1);
        for (auto e : ev_solved) cudaEventDestroy(e);
        for (auto e : ev_hdone) cudaEventDestroy(e);
        for (auto s : solve_streams) cudaStreamDestroy(s);
        if (stream3) cudaStreamDestroy(stream3);
        if (stream4) cudaStreamDestroy(stream4);
        for (auto& f : fan) if (f) cudaStreamDestroy(f);
        for (auto& e : ev_fan) if (e) cudaEventDestroy(e);
        if (ev_ab) cudaEventDestroy(ev_ab);
        if (ev_prod) cudaEventDestroy(ev_prod);
        if (stream2) cudaStreamDestroy(stream2);
        if (stream) cudaStreamDestroy(stream);
    }
    size_t proof_bytes() const { return n_commit ? 196 : 164; }
    uint64_t nZ_buckets_total(size_t n) const { return (uint64_t)n << (qZ.c - 1); }   // bucket sets of the Z query: one per proof
};

static inline int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    return atoi(v);
}

// table[w][i] = 2^(c w) * P_i for every query, so each MSM needs a single bucket set per proof
static void ctx_build_tables(Ctx& cx) {
    if (cx.tables_ready) return;
    cudaStream_t st = cx.stream;
    auto nwin = [](int c) { return (254 + c - 1) / c; };
    if (cx.qA.table.n) {   // rebuild after the Pedersen bases arrived: only those are missing
        if (cx.qPed.n && !cx.qPed.table.n) {
            cx.qPed.table.alloc((size_t)cx.qPed.n * nwin(cx.qPed.c));
            cx.qPedSigma.table.alloc((size_t)cx.qPedSigma.n * nwin(cx.qPedSigma.c));
            msm_precompute_g1(cx.ped_basis.p, cx.qPed.n, nwin(cx.qPed.c), cx.qPed.c, cx.qPed.table.p, st);
            msm_precompute_g1(cx.ped_basis_sigma.p, cx.qPedSigma.n, nwin(cx.qPedSigma.c), cx.qPedSigma.c, cx.qPedSigma.table.p, st);
            G16_CUDA(cudaStreamSynchronize(st));
        }
        cx.tables_ready = true;
        return;
    }
    cx.qA.table.alloc((size_t)cx.nA * nwin(cx.qA.c));
    cx.qB.table.alloc((size_t)cx.nB * nwin(cx.qB.c));
    cx.qZ.table.alloc((size_t)cx.nZ * nwin(cx.qZ.c));
    cx.qK.table.alloc((size_t)cx.nK * nwin(cx.qK.c));
    cx.tabB2.alloc((size_t)cx.nB2 * nwin(cx.cB2));
    msm_precompute_g1(cx.A.p, cx.nA, nwin(cx.qA.c), cx.qA.c, cx.qA.table.p, st);
    msm_precompute_g1(cx.B.p, cx.nB, nwin(cx.qB.c), cx.qB.c, cx.qB.table.p, st);
    msm_precompute_g1(cx.Z.p, cx.nZ, nwin(cx.qZ.c), cx.qZ.c, cx.qZ.table.p, st);
    msm_precompute_g1(cx.K.p, cx.nK, nwin(cx.qK.c), cx.qK.c, cx.qK.table.p, st);
    msm_precompute_g2(cx.B2.p, cx.nB2, nwin(cx.cB2), cx.cB2, cx.tabB2.p, st);
    if (cx.qPed.n) {
        cx.qPed.table.alloc((size_t)cx.qPed.n * nwin(cx.qPed.c));
        cx.qPedSigma.table.alloc((size_t)cx.qPedSigma.n * nwin(cx.qPedSigma.c));
        msm_precompute_g1(cx.ped_basis.p, cx.qPed.n, nwin(cx.qPed.c), cx.qPed.c, cx.qPed.table.p, st);
        msm_precompute_g1(cx.ped_basis_sigma.p, cx.qPedSigma.n, nwin(cx.qPedSigma.c), cx.qPedSigma.c, cx.qPedSigma.table.p, st);
    }
    G16_CUDA(cudaStreamSynchronize(st));
    cx.tables_ready = true;
}

// Evaluation-basis tables of the Z query. h = compute_h(a, b, c) is linear in (d, c): d_j = A(g w^j) B(g w^j) on the coset
// and c_j the evaluations of C on the domain (compute_h_run: h = M1 d - M2 c). Hence
//     Sum_k h_k Z_k = Sum_j d_j Qd_j + Sum_j c_j Qc_j,   Qd_j = Sum_k M1[k][j] Z_k,   Qc_j = -Sum_k M2[k][j] Z_k,
// the same group element as gnark's MultiExp(pk.G1.Z, h) (prove.go:267-275), so the proof bytes do not change. With these
// bases the prover skips the last two of the six transforms (coefficients of C, coefficients of E) and the subtraction;
// the evaluations of C are {0, +-1} and a few 34-bit sums in the ChaCha circuit, so the second sum is
// a handful of additions. Built once per context, as a DFT over group elements (default) or, as a cross-check
// (G16_EVAL_BUILD_MSM=1), column by column from the very transform compute_h_run uses (unit vectors through ntt_run, 512 at
// a time, each Q one row of the batched fixed-base MSM over pk.G1.Z).
static void ctx_build_eval_tables(Ctx& cx) {
    if (cx.eval_ready) return;
    cudaStream_t st = cx.stream;
    auto nwin = [](int c) { return (254 + c - 1) / c; };
    const uint32_t n = (uint32_t)cx.n_dom;
    const uint32_t blk = 512;
    const bool by_msm = env_int("G16_EVAL_BUILD_MSM", 0) != 0;   // cross-check builder: one batched MSM row per table point
    DevBuf<Fr> cols(by_msm ? (size_t)blk * n : 0);
    DevBuf<G1XYZZ> qx(n);
    DevBuf<G1Affine> qaff(n);
    for (int which = 0; which < 2; which++) {
        // columns of C beyond the last constraint only ever meet zero scalars: not built
        const uint32_t ncols = which ? cx.n_constraints : n;
        if (!by_msm) {
            // Q = DFT over the group of the pre-scaled key points (k_msm_g1.cu group_dft_g1): 15 stages of 16 384 scalar
            // products instead of 56 k MSM rows (7.7 s -> tens of ms)
            group_dft_g1(cx.Z.p, cx.nZ, cx.k_dom, which ? cx.dom.scale_ninv_den.p : cx.dom.scale_coset_inv_den.p, which,
                         cx.dom.tw_inv.p, qx.p, ncols, qaff.p, st);
        } else {
            G16_CUDA(cudaMemsetAsync(qx.p, 0, (size_t)n * sizeof(G1XYZZ), st));
            for (uint32_t j0 = 0; j0 < ncols; j0 += blk) {
                uint32_t rows = ncols - j0 < blk ? ncols - j0 : blk;
                compute_h_columns(cx.dom, cols.p, rows, j0, which, st);
                MsmShape sh = msm_make_shape(cx.qZ.n, rows, cx.qZ.c, 1);
                msm_run_g1(cx.ws1, sh, cx.qZ.table.p, cols.p, n, 1, nullptr, 1, st, nullptr);
                G16_CUDA(cudaMemcpyAsync(qx.p + j0, cx.ws1.result.p, (size_t)rows * sizeof(G1XYZZ), cudaMemcpyDeviceToDevice, st));
            }
            xyzz_to_affine_g1(qx.p, n, qaff.p, st);
        }
        PrecompQuery& q = which ? cx.qQc : cx.qQd;
        // The C half meets ~6 k scalars +-1 and a few hundred 34-bit values per proof: a narrow window keeps its bucket set
        // (and the reduction tree over it, which costs per bucket, not per entry) tiny. Measured with c = 13: 6 ms per
        // 1024 proofs in msm_tree_kernel alone, as much as the two transforms this path removes.
        q.n = ncols;
        q.c = which ? env_int("G16_C_QC", 8) : cx.qZ.c;
        q.table.alloc((size_t)q.n * nwin(q.c));
        msm_precompute_g1(qaff.p, q.n, nwin(q.c), q.c, q.table.p, st);
    }
    G16_CUDA(cudaStreamSynchronize(st));
    cx.ws1.log_reset();
    cx.eval_ready = true;
}

static std::unique_ptr<Ctx> ctx_create(const uint8_t* pk_bytes, size_t pk_len, const uint8_t* r1cs_bytes, size_t r1cs_len,
                                       int device) {
    std::unique_ptr<Ctx> cx(new Ctx());
    cx->device = device;
    G16_CUDA(cudaSetDevice(device));
    G16_CUDA(cudaStreamCreate(&cx->stream));
#if defined(G16_EMU)
    G16_CUDA(cudaStreamCreate(&cx->stream2));
#else
    {
        // The side stream carries the short wire-driven queries. At default priority its blocks are only dispatched when a
        // long kernel of the main stream drains, so the side stream trails the main one and the assembly waits for it;
        // at the highest priority its blocks take the next free slots and the work hides inside the long kernels.
        int least = 0, greatest = 0;
        G16_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        G16_CUDA(cudaStreamCreateWithPriority(&cx->stream2, cudaStreamDefault, env_int("G16_SIDE_PRIO", 1) ? greatest : least));
    }
#endif
    G16_CUDA(cudaEventCreate(&cx->ev_fork));
    G16_CUDA(cudaEventCreate(&cx->ev_join));
    G16_CUDA(cudaStreamCreate(&cx->stream3));
    G16_CUDA(cudaStreamCreate(&cx->stream4));
    G16_CUDA(cudaEventCreate(&cx->ev_join3));
    G16_CUDA(cudaEventCreate(&cx->ev_ab));
    G16_CUDA(cudaEventCreate(&cx->ev_prod));
    G16_CUDA(cudaEventCreate(&cx->ev_t0));
    G16_CUDA(cudaEventCreate(&cx->ev_t1));
    cudaStream_t st = cx->stream;
    // Round 1, measured on B200 (1024 ChaCha proofs): single stream 189.9 ms; two lanes 191.3 / 198.6 / 214.9 ms at sub-batch
    // 512 / 256 / 128 — the bucket-accumulation and NTT kernels filled every SM's register file, so kernels of the other
    // lane only time-sliced with them. End of round 2 (evaluation-basis Z query in both lanes, lighter kernels): two lanes
    // 125.9 ms against 128.9 ms on one main stream (sub-batch 384 / 342 / 256: 129.2 / 129.1 / 131.2), so circuits without a
    // commitment now default to the two lanes (G16_PIPELINE=0 / 1 forces; the stage timers only exist on the single stream,
    // g16_set_schedule switches per context).
    cx->pipeline = env_int("G16_PIPELINE", -1);
    cx->sub_batch = (uint32_t)env_int("G16_SUBBATCH", 512);
    cx->eval_z = env_int("G16_EVAL_Z", -1);
    cx->eval_z_min = (uint32_t)env_int("G16_EVAL_Z_MIN", 128);
    if (cx->sub_batch == 0) cx->sub_batch = 1;
    cx->pipeline_stagger = env_int("G16_PIPE_STAGGER", 1) != 0;
    cx->test_lane_oom = env_int("G16_TEST_LANE_OOM", 0) != 0;
    cx->solve_chains = (size_t)env_int("G16_SOLVE_CHAINS", 2);
    if (cx->solve_chains < 2 || cx->solve_chains > 16) cx->solve_chains = 2;
    cx->split_solve = env_int("G16_SPLIT_SOLVE", -1);   // -1: decided once the circuit is known (below)
    cx->bitq_state = env_int("G16_BITQ", 1) ? 0 : 2;
    cx->bitq_min_rows = (uint32_t)env_int("G16_BITQ_MIN_ROWS", 256);
    cx->bitq_min_batch = (uint32_t)env_int("G16_BITQ_MIN_BATCH", 32);
    cx->bitq_test_exception = env_int("G16_BITQ_TEST_EXC", 0) != 0;

    PkFile pk = parse_pk(pk_bytes, pk_len);
    R1csFile cs = parse_r1cs(r1cs_bytes, r1cs_len);
    if (pk.nb_wires != cs.n_wires()) throw ParseError("pk and r1cs disagree on the number of wires");
    if (pk.n_commit_keys != cs.commitments.size()) throw ParseError("pk and r1cs disagree on the number of commitments");
    uint64_t need = 1;
    while (need < cs.n_constraints) need <<= 1;
    if (need != pk.n) throw ParseError("pk domain size does not match the constraint count");
    cx->n_dom = pk.n;
    while ((1ull << cx->k_dom) < pk.n) cx->k_dom++;
    cx->nA = pk.nA; cx->nB = pk.nB; cx->nZ = pk.nZ; cx->nK = pk.nK; cx->nB2 = pk.nB2;
    cx->nb_wires = (uint32_t)pk.nb_wires;
    cx->n_public = (uint32_t)cs.n_public; cx->n_secret = (uint32_t)cs.n_secret;
    cx->n_constraints = (uint32_t)cs.n_constraints;
    cx->n_instr = (uint32_t)cs.n_instr(); cx->nlevels = (uint32_t)cs.levels.size();
    cx->n_commit = (uint32_t)cs.commitments.size();
    // concurrent solves of the sub-batches (135.0 -> 134.4 ms per 1024 ChaCha proofs); circuits with a commitment keep the
    // single chain (their solve includes a Pedersen MSM per sub-batch)
    if (cx->split_solve < 0) cx->split_solve = cx->n_commit ? 0 : 2;
    if (cx->pipeline < 0) cx->pipeline = cx->n_commit ? 0 : 1;

    // ---- decompress the key on the GPU (~89k Fp + 12.5k Fp2 square roots)
    DevBuf<uint32_t> err(1);
    err.zero(st);
    auto decompress1 = [&](const uint8_t* raw, uint32_t n, DevBuf<G1Affine>& out) {
        DevBuf<uint8_t> d;
        d.upload(raw, (size_t)n * 32, st);
        out.alloc(n);
        launch_decompress_g1(d.p, n, out.p, err.p, st);
        G16_CUDA(cudaStreamSynchronize(st));
    };
    auto decompress2 = [&](const uint8_t* raw, uint32_t n, DevBuf<G2Affine>& out) {
        DevBuf<uint8_t> d;
        d.upload(raw, (size_t)n * 64, st);
        out.alloc(n);
        launch_decompress_g2(d.p, n, out.p, err.p, st);
        G16_CUDA(cudaStreamSynchronize(st));
    };
    DevBuf<G1Affine> abd;
    DevBuf<G2Affine> bd2;
    decompress1(pk.g1_abd, 3, abd);
    decompress1(pk.A, pk.nA, cx->A);
    decompress1(pk.B, pk.nB, cx->B);
    decompress1(pk.Z, pk.nZ, cx->Z);
    decompress1(pk.K, pk.nK, cx->K);
    decompress2(pk.g2_bd, 2, bd2);
    decompress2(pk.B2, pk.nB2, cx->B2);
    uint32_t herr = 0;
    err.download(&herr, 1, st);
    G16_CUDA(cudaStreamSynchronize(st));
    if (herr) throw ParseError("pk: point decompression failed (flags=" + std::to_string(herr) + ")");
    {
        // gnark's decoder (ProvingKey.ReadFrom -> G2Affine.SetBytes) also checks that G2 points lie in the r-torsion subgroup
        PairingWorkspace pw;
        pairing_consts_ensure(pw, st);
        DevBuf<G2Affine> all(2 + (size_t)pk.nB2);
        G16_CUDA(cudaMemcpyAsync(all.p, bd2.p, 2 * sizeof(G2Affine), cudaMemcpyDeviceToDevice, st));
        if (pk.nB2) G16_CUDA(cudaMemcpyAsync(all.p + 2, cx->B2.p, (size_t)pk.nB2 * sizeof(G2Affine), cudaMemcpyDeviceToDevice, st));
        DevBuf<uint8_t> okb(all.n);
        launch_g2_subgroup(all.p, (uint32_t)all.n, pw.consts.p, okb.p, st);
        std::vector<uint8_t> hok(all.n);
        okb.download(hok.data(), all.n, st);
        G16_CUDA(cudaStreamSynchronize(st));
        for (uint8_t v : hok)
            if (!v) throw ParseError("pk: a G2 point is not in the correct subgroup");
    }
    {
        G1Affine h1[3];
        G2Affine h2[2];
        abd.download(h1, 3, st);
        bd2.download(h2, 2, st);
        G16_CUDA(cudaStreamSynchronize(st));
        cx->keys.alpha = h1[0]; cx->keys.beta = h1[1]; cx->keys.delta = h1[2];
        cx->keys.beta2 = h2[0]; cx->keys.delta2 = h2[1];
        cx->delta_tab.alloc(64 * 15);
        cx->delta2_tab.alloc(64 * 15);
        cx->keys.delta_tab = cx->delta_tab.p;
        cx->keys.delta2_tab = cx->delta2_tab.p;
        launch_fixed_base_tables(cx->keys, cx->delta_tab.p, cx->delta2_tab.p, st);
        G16_CUDA(cudaStreamSynchronize(st));
    }

    // ---- wire maps of the four G1 queries (the G2 query shares B's)
    std::vector<uint32_t> mapA, mapB, mapK;
    for (uint32_t w = 0; w < cx->nb_wires; w++) {
        if (!pk.inf_a[w]) mapA.push_back(w);
        if (!pk.inf_b[w]) mapB.push_back(w);
    }
    {
        std::vector<uint8_t> skip(cx->nb_wires, 0);
        for (auto& ci : cs.commitments) {
            for (uint32_t w : ci.private_committed) if (w < cx->nb_wires) skip[w] = 1;
            if (ci.commitment_index < cx->nb_wires) skip[ci.commitment_index] = 1;
        }
        for (uint32_t w = cx->n_public; w < cx->nb_wires; w++) if (!skip[w]) mapK.push_back(w);
    }
    if (mapA.size() != pk.nA || mapB.size() != pk.nB) throw ParseError("pk: infinity masks do not match the query sizes");
    if (mapK.size() != pk.nK) throw ParseError("pk: len(G1.K) does not match the private wires of the r1cs");
    cx->h_mapA = mapA; cx->h_mapB = mapB; cx->h_mapK = mapK;
    cx->qA.map.upload(mapA.data(), mapA.size(), st);
    cx->qB.map.upload(mapB.data(), mapB.size(), st);
    cx->qK.map.upload(mapK.data(), mapK.size(), st);
    G16_CUDA(cudaStreamSynchronize(st));

    // ---- fixed-base tables (built now, or on the first prove when G16_LAZY_TABLES=1 — gnark's icicle backend also
    //      defers its device set-up to the first Prove)
    // Z query: c = 15 for the 2^15 domain of ChaCha (14 / 16 measured: 153.1 / 158.0 vs 150.5 ms per 1024 proofs); the 2^17
    // domain of the AES circuits has four times the entries per bucket, so one window less pays for the larger tree
    // (AES-128, 256 proofs: 171.5 -> 168.8 ms at c = 16, 188.0 at c = 14; profiles/sweep_r02g_aes_windows.jsonl)
    cx->qZ.c = env_int("G16_C_Z", pk.nZ >= (1u << 16) ? 16 : 15);
    cx->qA.c = env_int("G16_C_A", 13);
    cx->qB.c = env_int("G16_C_B", 13);
    cx->qK.c = env_int("G16_C_K", 13);
    cx->cB2 = env_int("G16_C_B2", 13);
    cx->qA.n = pk.nA; cx->qB.n = pk.nB; cx->qZ.n = pk.nZ; cx->qK.n = pk.nK;
    if (!env_int("G16_LAZY_TABLES", 0)) ctx_build_tables(*cx);

    // ---- FFT domain from the pk header (w, g big-endian canonical -> Montgomery on the device)
    {
        DevBuf<uint8_t> hb;
        hb.upload(&pk.fr_hdr[0][0], 160, st);
        DevBuf<Fr> hm(5);
        fr_be_to_mont(hb.p, 5, hm.p, st);
        Fr h[5];
        hm.download(h, 5, st);
        G16_CUDA(cudaStreamSynchronize(st));
        ntt_domain_init(cx->dom, cx->k_dom, h[1], h[3], st);
    }

    // ---- solver program: static analysis of which wire every R1C defines (instruction order is a valid schedule,
    //      SURVEY.md Appendix E), then upload
    {
        std::vector<uint8_t> solved(cx->nb_wires, 0);
        for (uint32_t w = 0; w < cx->n_public + cx->n_secret; w++) solved[w] = 1;
        std::vector<InsMeta> meta(cs.n_instr());
        std::vector<uint32_t> lk_index(cs.bp_kind.size(), 0);
        uint32_t ntab = 0;
        for (size_t b = 0; b < cs.bp_kind.size(); b++) if (cs.bp_kind[b] == INS_LOOKUP) lk_index[b] = ntab++;
        const std::vector<uint32_t>& cd = cs.calldata;
        std::vector<uint32_t> count_ids;
        uint32_t count_words = 0;
        for (size_t i = 0; i < cs.n_instr(); i++) {
            InsMeta& m = meta[i];
            uint64_t s0 = cs.start[i];
            if (s0 >= cd.size() || s0 + cd[s0] > cd.size()) throw ParseError("r1cs: instruction calldata out of range");
            m.cd_start = (uint32_t)s0;
            uint8_t kind = cs.bp_kind[cs.bp_id[i]];
            m.kind = kind;
            m.solve_wire = SOLVE_WIRE_NONE;
            m.cons_off = cs.cons_off[i];
            m.wire_off = cs.wire_off[i];
            m.lookup_tab = lk_index[cs.bp_id[i]];
            if (kind == INS_R1C) {
                uint32_t n[3] = {cd[s0 + 1], cd[s0 + 2], cd[s0 + 3]};
                size_t pos = s0 + 4;
                int uside = -1;
                uint32_t uw = SOLVE_WIRE_NONE;
                for (int side = 0; side < 3; side++)
                    for (uint32_t t = 0; t < n[side]; t++) {
                        uint32_t wid = cd[pos + 1];
                        pos += 2;
                        if (wid == WIRE_CONST) continue;
                        if (wid >= cx->nb_wires) throw ParseError("r1cs: wire id out of range");
                        if (!solved[wid]) {
                            if (uside >= 0 && (uside != side || uw != wid)) throw ParseError("r1cs: constraint with two unknown wires");
                            uside = side; uw = wid;
                        }
                    }
                if (m.cons_off >= cx->n_constraints) throw ParseError("r1cs: constraint offset out of range");
                if (uside >= 0) { m.solve_wire = uw; m.kind |= (uint32_t)uside << 8; solved[uw] = 1; }
            } else if (kind == INS_HINT) {
                uint32_t hid = cd[s0 + 1], nin = cd[s0 + 2];
                size_t pos = s0 + 3;
                for (uint32_t k = 0; k < nin; k++) { uint32_t nt = cd[pos]; pos += 1 + 2 * (size_t)nt; }
                uint32_t o0 = cd[pos], o1 = cd[pos + 1];
                if (o1 < o0 || o1 > cx->nb_wires) throw ParseError("r1cs: hint output range out of bounds");
                for (uint32_t w = o0; w < o1; w++) solved[w] = 1;
                if (hid == HINT_COUNT) {   // room for the query index built on the device below
                    m.lookup_tab = count_words;
                    count_ids.push_back((uint32_t)i);
                    count_words += 6 + nin / 2;
                }
                if (hid == HINT_RANDOMIZE) cx->has_randomize = true;
                else if (hid == HINT_BSB22) {
                    if (cx->bsb_ins != 0xFFFFFFFFu) { cx->solver_supported = false; cx->solver_unsupported_reason = "more than one BSB22 commitment"; }
                    cx->bsb_ins = (uint32_t)i;
                    cx->commit_wire = o0;
                } else if (hid != HINT_NBITS && hid != HINT_COUNT) {
                    cx->solver_supported = false;
                    cx->solver_unsupported_reason = "hint id " + std::to_string(hid) + " is not implemented on the device";
                }
            } else {
                uint32_t nin = cd[s0 + 2];
                if (m.wire_off + nin > cx->nb_wires) throw ParseError("r1cs: lookup output range out of bounds");
                for (uint32_t k = 0; k < nin; k++) solved[m.wire_off + k] = 1;
            }
        }
        for (uint32_t w = 0; w < cx->nb_wires; w++) if (!solved[w]) throw ParseError("r1cs: wire " + std::to_string(w) + " is never defined");
        // Inside a level the instructions are independent, so they may be reordered: those whose linear expressions are
        // long (the 33-bit adder recompositions of ChaCha: ~130 terms) are moved to the end of their level and solved by the
        // term-parallel kernel (a warp per instruction and witness), the short ones by the witness-parallel kernel. The
        // level time is then bounded by a ~16-term chain instead of a 130-term one.
        const uint32_t long_terms = (uint32_t)env_int("G16_SOLVER_LONG", 24);
        auto n_terms = [&](uint32_t id) -> uint32_t {
            uint64_t s0 = cs.start[id];
            uint8_t kind = cs.bp_kind[cs.bp_id[id]];
            if (kind == INS_R1C) return cd[s0 + 1] + cd[s0 + 2] + cd[s0 + 3];
            if (kind == INS_HINT && cd[s0 + 1] == HINT_NBITS && cd[s0 + 2] == 1) return cd[s0 + 3];
            return 0;   // lookups and the other hints stay on the witness-parallel kernel
        };
        auto is_count = [&](uint32_t id) -> bool {
            uint64_t s0 = cs.start[id];
            return cs.bp_kind[cs.bp_id[id]] == INS_HINT && cd[s0 + 1] == HINT_COUNT;
        };
        std::vector<uint32_t> lvl_instr, lvl_off(1, 0), lvl_split, lvl_split2;
        for (auto& lv : cs.levels) {
            std::vector<uint32_t> longs, counts;
            for (uint32_t id : lv) {
                if (id >= cs.n_instr()) throw ParseError("r1cs: level references unknown instruction");
                if (id == cx->bsb_ins) cx->bsb_level = (uint32_t)(lvl_off.size() - 1);
                if (is_count(id)) counts.push_back(id);   // logderivarg.countHint: a block per instruction and witness
                else if (n_terms(id) > long_terms) longs.push_back(id);
                else lvl_instr.push_back(id);
            }
            lvl_split.push_back((uint32_t)lvl_instr.size());
            lvl_instr.insert(lvl_instr.end(), longs.begin(), longs.end());
            lvl_split2.push_back((uint32_t)lvl_instr.size());
            lvl_instr.insert(lvl_instr.end(), counts.begin(), counts.end());
            lvl_off.push_back((uint32_t)lvl_instr.size());
        }
        cx->h_level_split = lvl_split;
        cx->h_level_split2 = lvl_split2;
        if (lvl_instr.size() != cs.n_instr()) throw ParseError("r1cs: levels do not cover every instruction exactly once");
        // lookup tables: entry k of table t = coeffs[cid] of the single constant term of that entry (pure gather)
        std::vector<uint64_t> tabs((size_t)(ntab ? ntab : 1) * 256 * 4, 0);
        for (size_t b = 0; b < cs.bp_kind.size(); b++) {
            if (cs.bp_kind[b] != INS_LOOKUP) continue;
            const auto& ec = cs.bp_lookup_entries[b];
            size_t q = 0, ent = 0;
            while (q < ec.size() && ent < 256) {
                uint32_t nt = ec[q++];
                if (nt != 1 || q + 2 > ec.size() || ec[q + 1] != WIRE_CONST || ec[q] * 4ull + 4 > cs.coeffs.size())
                    throw ParseError("r1cs: unsupported lookup table entry");
                memcpy(&tabs[((size_t)lk_index[b] * 256 + ent) * 4], &cs.coeffs[(size_t)ec[q] * 4], 32);
                q += 2;
                ent++;
            }
        }
        cx->d_calldata.upload(cd.data(), cd.size(), st);
        cx->d_meta.upload(meta.data(), meta.size(), st);
        cx->d_level_instr.upload(lvl_instr.data(), lvl_instr.size(), st);
        cx->d_level_off.upload(lvl_off.data(), lvl_off.size(), st);
        cx->h_level_off = lvl_off;
        cx->d_coeffs.upload((const Fr*)cs.coeffs.data(), cs.coeffs.size() / 4, st);
        cx->d_lookup_tabs.upload((const Fr*)tabs.data(), tabs.size() / 4, st);
        cx->d_ucoef_inv.alloc(cs.n_instr());
        SolverProgram& sp = cx->sp;
        sp.calldata = cx->d_calldata.p; sp.meta = cx->d_meta.p; sp.level_instr = cx->d_level_instr.p;
        sp.level_off = cx->d_level_off.p; sp.coeffs = cx->d_coeffs.p; sp.ucoef_inv = cx->d_ucoef_inv.p;
        sp.lookup_tabs = cx->d_lookup_tabs.p; sp.nlevels = cx->nlevels; sp.n_wires = cx->nb_wires;
        sp.n_dom = (uint32_t)cx->n_dom; sp.fast_coeffs = 0;
        sp.randomize = nullptr; sp.bsb_ins = cx->bsb_ins;
        sp.count_index = nullptr;
        if (!count_ids.empty()) {
            DevBuf<uint32_t> d_ids;
            d_ids.upload(count_ids.data(), count_ids.size(), st);
            cx->d_count_index.alloc(count_words);
            sp.count_index = cx->d_count_index.p;
            launch_solver_count_index(sp, d_ids.p, (uint32_t)count_ids.size(), cx->d_count_index.p, st);
            G16_CUDA(cudaStreamSynchronize(st));
        }
        G16_CUDA(cudaStreamSynchronize(st));   // host vectors above must outlive the copies
        sp.fast_coeffs = launch_solver_init(sp, (uint32_t)cs.n_instr(), (uint32_t)(cs.coeffs.size() / 4), cx->d_ucoef_inv.p, st);
    }
    if (cx->n_commit > 1 || (cx->n_commit == 1 && (cx->bsb_ins == 0xFFFFFFFFu || pk.ped.size() != 1))) {
        cx->solver_supported = false;
        if (cx->solver_unsupported_reason.empty()) cx->solver_unsupported_reason = "only circuits with at most one BSB22 commitment are supported";
    }
    if (cx->n_commit == 1 && cx->solver_supported) {
        const CommitmentInfo& ci = cs.commitments[0];
        if (ci.commitment_index != cx->commit_wire) throw ParseError("r1cs: commitment wire does not match the Bsb22 hint output");
        if (ci.private_committed.size() != pk.ped[0].n_basis) throw ParseError("pk: Pedersen basis size does not match the committed wires");
        if (!ci.public_and_commitment_committed.empty()) throw ParseError("public-committed wires are not supported");
        DevBuf<G1Affine> basis, basis_sigma;
        decompress1(pk.ped[0].basis, pk.ped[0].n_basis, basis);
        decompress1(pk.ped[0].basis_sigma, pk.ped[0].n_sigma, basis_sigma);
        uint32_t herr2 = 0;
        err.download(&herr2, 1, st);
        G16_CUDA(cudaStreamSynchronize(st));
        if (herr2) throw ParseError("pk: Pedersen basis decompression failed");
        cx->ped_basis = std::move(basis);
        cx->ped_basis_sigma = std::move(basis_sigma);
        cx->qPed.n = cx->qPedSigma.n = pk.ped[0].n_basis;
        // c = 11: the ~6 k committed wires of the AES circuits give ~30 k entries per proof, so a bucket set of 2^15 (c = 16, the
        // value until late round 2) cost 9.4 ms of reduction tree per 256 proofs and query for 0.65 ms of accumulation
        // (profiles/launches_r02g_aes128_batch256_summary.txt); sweep 10..16 in profiles/sweep_r02g_aes_pedersen_window.log
        cx->qPed.c = cx->qPedSigma.c = env_int("G16_C_PED", 11);
        cx->qPed.map.upload(ci.private_committed.data(), ci.private_committed.size(), st);
        cx->qPedSigma.map.upload(ci.private_committed.data(), ci.private_committed.size(), st);
        G16_CUDA(cudaStreamSynchronize(st));
        if (cx->tables_ready) { cx->tables_ready = false; ctx_build_tables(*cx); }
    }
    cx->d_status.alloc(1);
    cx->h_status.assign(1, 0);
    cx->solver_graphs = solver_graph_cache_create();
    return cx;
}

// ---------------------------------------------------------------------------------------------------------------------
// batch pipeline. Inputs must already be on the device: either the ChaCha request arrays (d_keys, ...) or d_witness.
// ---------------------------------------------------------------------------------------------------------------------
static void ctx_ensure_batch(Ctx& cx, size_t n) {
    cx.W.ensure(n * cx.nb_wires);
    cx.Aev.ensure(n * cx.n_dom);
    cx.Bev.ensure(n * cx.n_dom);
    cx.Cev.ensure(n * cx.n_dom);
    cx.resA.ensure(n); cx.resB1.ensure(n); cx.resK.ensure(n); cx.resZ.ensure(n); cx.resB2.ensure(n); cx.resZc.ensure(n);
    cx.d_proofs.ensure(n * cx.proof_bytes());
    cx.d_ct.ensure(n * 64);
    cx.d_rs.ensure(2 * n);
    cx.d_status.ensure(n);
    if (cx.n_commit) { cx.resCommit.ensure(n); cx.resPok.ensure(n); cx.commit_aff.ensure(n); }
}

static void run_query_g1(MsmWorkspace<G1>& ws, cudaStream_t st, const PrecompQuery& q, const Fr* scalars, size_t row_stride,
                         size_t elem_stride, bool use_map, uint32_t rows, G1XYZZ* out, StageTimer* tm) {
    MsmShape sh = msm_make_shape(q.n, rows, q.c, 1);
    msm_run_g1(ws, sh, q.table.p, scalars, row_stride, elem_stride, use_map ? q.map.p : nullptr, 1, st, tm);
    G16_CUDA(cudaMemcpyAsync(out, ws.result.p, (size_t)rows * sizeof(G1XYZZ), cudaMemcpyDeviceToDevice, st));
}

// The batch-affine levels in front of the Z query's accumulation pay from ~16 ChaCha proofs on: below 2^23 entries their
// fixed latency (nine more launches, three rounds of inversions) is larger than the products they save, and since the
// wire-driven queries of a small batch no longer hide it (ctx_wire_queries_fan) it is on the critical path — 4 proofs 4.18 ->
// 3.40 ms, 8 proofs 4.89 -> 4.73 ms, 16 and more unchanged (profiles/sweep_r02j_query_fan.log). G16_Z_BA_MIN moves the threshold.
static void z_query_ba_policy(MsmWorkspace<G1>& ws, const PrecompQuery& q, uint32_t rows) {
    static const int zmin = env_int("G16_Z_BA_MIN", 1 << 23);
    static const bool forced = getenv("G16_MSM_BA_MIN") != nullptr;   // an explicit MSM threshold (the tests' forced variants) decides alone
    if (forced) { ws.no_ba = false; return; }
    const size_t entries = (size_t)rows * q.n * (size_t)((254 + q.c - 1) / q.c);
    ws.no_ba = entries < (size_t)zmin;
}

// ---- combination tables (k_bitq.cu)
static void bitq_build_one(Ctx& cx, BitQuery& bq, const PrecompQuery& q, const std::vector<uint32_t>& map, const G1Affine* pts1,
                           const G2Affine* pts2, const G2Affine* tab2, int c2, cudaStream_t st) {
    auto nwin = [](int c) { return (254 + c - 1) / c; };
    std::vector<uint32_t> bitpts, tripts, rest;
    for (uint32_t i = 0; i < map.size(); i++) {
        const uint32_t m = cx.bit_mask[map[i]];
        ((m & 1u) ? bitpts : ((m & 2u) ? tripts : rest)).push_back(i);
    }
    // worth it only when most of the query is bits / trits (AES: bytes and field elements, the general path stays)
    bq.on = bitpts.size() + tripts.size() >= BITQ_K && 2 * (bitpts.size() + tripts.size()) >= map.size();
    if (!bq.on) return;
    bq.groups_bin = (uint32_t)((bitpts.size() + BITQ_K - 1) / BITQ_K);
    bq.groups = bq.groups_bin + (uint32_t)((tripts.size() + BITQ_T - 1) / BITQ_T);
    std::vector<uint32_t> gw((size_t)bq.groups * BITQ_K, BITQ_NONE), gp((size_t)bq.groups * BITQ_K, BITQ_NONE);
    for (size_t k = 0; k < bitpts.size(); k++) { gp[k] = bitpts[k]; gw[k] = map[bitpts[k]]; }
    for (size_t k = 0; k < tripts.size(); k++) {   // five used slots of the eight of a ternary group
        const size_t slot = ((size_t)bq.groups_bin + k / BITQ_T) * BITQ_K + k % BITQ_T;
        gp[slot] = tripts[k]; gw[slot] = map[tripts[k]];
    }
    DevBuf<uint32_t> d_gp, d_rest;
    d_gp.upload(gp.data(), gp.size(), st);
    bq.grp_wires.upload(gw.data(), gw.size(), st);
    bq.table1.alloc((size_t)bq.groups << BITQ_K);
    bitq_build_g1(pts1, d_gp.p, bq.groups, bq.groups_bin, bq.table1.p, st);
    if (pts2) {
        bq.table2.alloc((size_t)bq.groups << BITQ_K);
        bitq_build_g2(pts2, d_gp.p, bq.groups, bq.groups_bin, bq.table2.p, st);
    }
    bq.rest.n = (uint32_t)rest.size();
    bq.rest.c = q.c;
    if (!rest.empty()) {
        std::vector<uint32_t> rmap(rest.size());
        for (size_t k = 0; k < rest.size(); k++) rmap[k] = map[rest[k]];
        bq.rest.map.upload(rmap.data(), rmap.size(), st);
        d_rest.upload(rest.data(), rest.size(), st);
        bq.rest.table.alloc(rest.size() * nwin(q.c));
        bitq_gather_g1(q.table.p, q.n, nwin(q.c), d_rest.p, bq.rest.n, bq.rest.table.p, st);
        if (tab2) {
            bq.rest_tab2.alloc(rest.size() * nwin(c2));
            bitq_gather_g2(tab2, q.n, nwin(c2), d_rest.p, bq.rest.n, bq.rest_tab2.p, st);
        }
    }
    G16_CUDA(cudaStreamSynchronize(st));   // the host vectors and the temporaries above must outlive the copies
}
static void ctx_bitq_build(Ctx& cx) {
    if (cx.bitq_built) return;
    cudaStream_t st = cx.stream;
    bitq_build_one(cx, cx.bqA, cx.qA, cx.h_mapA, cx.A.p, nullptr, nullptr, 0, st);
    bitq_build_one(cx, cx.bqB, cx.qB, cx.h_mapB, cx.B.p, cx.B2.p, cx.tabB2.p, cx.cB2, st);
    bitq_build_one(cx, cx.bqK, cx.qK, cx.h_mapK, cx.K.p, nullptr, nullptr, 0, st);
    cx.d_bitq_exc.ensure(1);
    cx.bitq_built = true;
    if (!cx.bqA.on && !cx.bqB.on && !cx.bqK.on) cx.bitq_state = 2;   // not a bit-level circuit (AES): nothing to gain
}
// one G1 query through its combination table: general path over the non-bit points + one table point per group of bit wires
static void run_query_bitq_g1(Ctx& cx, cudaStream_t st, const BitQuery& bq, const Fr* w, size_t n, uint32_t rows, bool fresh_entries,
                              G1XYZZ* out) {
    if (bq.rest.n) run_query_g1(cx.ws1b, st, bq.rest, w, 1, n, true, rows, out, nullptr);
    else G16_CUDA(cudaMemsetAsync(out, 0, (size_t)rows * sizeof(G1XYZZ), st));
    cx.bitq_entries.ensure((size_t)rows * bq.groups);
    cx.bitq_tmp1.ensure(rows);
    if (fresh_entries) bitq_entries(w, n, rows, bq.grp_wires.p, bq.groups, bq.groups_bin, cx.bitq_entries.p, cx.d_bitq_exc.p, st);
    msm_sum_rows_g1(cx.ws1b, bq.table1.p, cx.bitq_entries.p, rows * bq.groups, rows, cx.bitq_tmp1.p, st);
    xyzz_add_g1(out, cx.bitq_tmp1.p, rows, st);
}

// Solves witnesses [sb, sb + rows) of the batch (W wire-major, stride n) on stream st. With a BSB22 commitment the level
// schedule is cut after the level that holds the commitment hint: Pedersen MSM over the committed wires -> hash to field ->
// challenge wire (gnark prove.go:84-108 overrides the hint the same way), then the remaining levels run. wsc: the MSM
// workspace the commitment may use on this stream.
static size_t ctx_solve(Ctx& cx, size_t n, size_t sb, uint32_t rows, cudaStream_t st, MsmWorkspace<G1>& wsc) {
    SolverProgram sp = cx.sp;
    sp.randomize = cx.has_randomize ? cx.d_mask.p + sb : nullptr;
    Fr* W = cx.W.p + sb;
    Fr* A = cx.Aev.p + sb * cx.n_dom;
    Fr* B = cx.Bev.p + sb * cx.n_dom;
    Fr* C = cx.Cev.p + sb * cx.n_dom;
    uint32_t* status = cx.d_status.p + sb;
    if (!cx.n_commit)
        return launch_solver(sp, cx.h_level_off.data(), cx.h_level_split.data(), cx.h_level_split2.data(), 0, cx.nlevels, rows, W, n, A, B, C, status, st, cx.solver_graphs);
    size_t launches = launch_solver(sp, cx.h_level_off.data(), cx.h_level_split.data(), cx.h_level_split2.data(), 0, cx.bsb_level + 1, rows, W, n, A, B, C, status, st, cx.solver_graphs);
    for (size_t o = 0; o < rows; o += cx.sub_batch) {
        uint32_t r = (uint32_t)((rows - o) < cx.sub_batch ? (rows - o) : cx.sub_batch);
        run_query_g1(wsc, st, cx.qPed, W + o, 1, n, true, r, cx.resCommit.p + sb + o, nullptr);
    }
    launch_bsb22_challenge(cx.resCommit.p + sb, rows, W, n, cx.commit_wire, cx.commit_aff.p + sb, st);
    launches += 1 + launch_solver(sp, cx.h_level_off.data(), cx.h_level_split.data(), cx.h_level_split2.data(), cx.bsb_level + 1, cx.nlevels, rows, W, n, A, B, C, status, st, cx.solver_graphs);
    return launches;
}

// Small batches (G16_QUERY_FAN, <= 4 proofs): every wire-driven query is a chain of ~15 short, latency-bound launches (digits,
// scans, accumulate, merge levels, tree), and one after the other on the side stream they took 1.8 ms for one request — 0.8 ms
// longer than the transforms and the Z query on the main stream, which then waited for them. Here A stays on the side stream and
// B1, B2 and K each get a stream and a workspace of their own; the side stream joins them again, and the half-products of the
// assembly (stream4) are released as soon as A and B1 are there.
static void ctx_wire_queries_fan(Ctx& cx, size_t n, uint32_t rows, cudaStream_t st2) {
    const Fr* w = cx.W.p;
    if (!cx.fan[0]) {
        for (auto& f : cx.fan) G16_CUDA(cudaStreamCreate(&f));
        for (auto& e : cx.ev_fan) G16_CUDA(cudaEventCreate(&e));
    }
    cx.ws1b.no_ba = cx.wsf[0].no_ba = cx.wsf[1].no_ba = cx.n_commit == 0;
    G16_CUDA(cudaEventRecord(cx.ev_fan[0], st2));   // the witness is complete (the side stream has waited for the solver)
    for (auto& f : cx.fan) G16_CUDA(cudaStreamWaitEvent(f, cx.ev_fan[0], 0));
    run_query_g1(cx.ws1b, st2, cx.qA, w, 1, n, true, rows, cx.resA.p, nullptr);
    run_query_g1(cx.wsf[0], cx.fan[0], cx.qB, w, 1, n, true, rows, cx.resB1.p, nullptr);
    G16_CUDA(cudaEventRecord(cx.ev_fan[1], cx.fan[0]));
    {
        MsmShape sh = msm_make_shape(cx.nB2, rows, cx.cB2, 1);
        msm_run_g2(cx.ws2, sh, cx.tabB2.p, w, 1, n, cx.qB.map.p, 1, cx.fan[1], nullptr);
        G16_CUDA(cudaMemcpyAsync(cx.resB2.p, cx.ws2.result.p, (size_t)rows * sizeof(G2XYZZ), cudaMemcpyDeviceToDevice, cx.fan[1]));
        G16_CUDA(cudaEventRecord(cx.ev_fan[2], cx.fan[1]));
    }
    run_query_g1(cx.wsf[1], cx.fan[2], cx.qK, w, 1, n, true, rows, cx.resK.p, nullptr);
    G16_CUDA(cudaEventRecord(cx.ev_fan[3], cx.fan[2]));
    G16_CUDA(cudaStreamWaitEvent(st2, cx.ev_fan[1], 0));
    G16_CUDA(cudaEventRecord(cx.ev_ab, st2));   // A and B1 of the whole batch are complete
    if (cx.n_commit)   // proof of knowledge of the commitment: same scalars over BasisExpSigma
        run_query_g1(cx.ws1b, st2, cx.qPedSigma, w, 1, n, true, rows, cx.resPok.p, nullptr);
    G16_CUDA(cudaStreamWaitEvent(st2, cx.ev_fan[2], 0));
    G16_CUDA(cudaStreamWaitEvent(st2, cx.ev_fan[3], 0));
}

// the wire-driven queries of one sub-batch (A, B1, K on G1, B on G2, the commitment PoK): short, latency-bound kernels
static void ctx_wire_queries(Ctx& cx, size_t n, size_t sb, uint32_t rows, cudaStream_t st2, bool eval_z = false) {
    const Fr* w = cx.W.p + sb;   // column offset into the wire-major array
    // The side stream's general-path queries meet almost only 0 / +-1 scalars: their live entries are a small fraction of the
    // upper bound the batch-affine decision is taken on, so the levels would be launches over mostly idle grids (measured: 9
    // launches, ~1 ms per sub-batch for the C-evaluation query alone).
    cx.ws1b.no_ba = cx.n_commit == 0;
    if (eval_z)   // C-evaluation half of the Z query: almost every scalar is 0 or +-1
        run_query_g1(cx.ws1b, st2, cx.qQc, cx.Cev.p + sb * cx.n_dom, cx.n_dom, 1, false, rows, cx.resZc.p + sb, nullptr);
    const bool bitq = cx.bitq_state == 1 && cx.bitq_built && n >= cx.bitq_min_batch;
    static const int fan_max = env_int("G16_QUERY_FAN", 4);   // measured: 1 / 2 proofs 3.49 / 3.72 -> 2.81 / 3.08 ms, 8 / 16 proofs 4.85 / 6.57 -> 4.95 / 6.60 ms
    if (!eval_z && !bitq && sb == 0 && rows == n && (int)n <= fan_max) {
        ctx_wire_queries_fan(cx, n, rows, st2);
        return;
    }
    if (bitq && cx.bqA.on) run_query_bitq_g1(cx, st2, cx.bqA, w, n, rows, true, cx.resA.p + sb);
    else run_query_g1(cx.ws1b, st2, cx.qA, w, 1, n, true, rows, cx.resA.p + sb, nullptr);
    if (bitq && cx.bqB.on) {
        const BitQuery& bq = cx.bqB;
        run_query_bitq_g1(cx, st2, bq, w, n, rows, true, cx.resB1.p + sb);
        if (sb + rows == n) G16_CUDA(cudaEventRecord(cx.ev_ab, st2));   // A and B1 of the whole batch are complete
        // the G2 element of the same wires: same entries, the G2 table
        G2XYZZ* out2 = cx.resB2.p + sb;
        if (bq.rest.n) {
            MsmShape sh = msm_make_shape(bq.rest.n, rows, cx.cB2, 1);
            msm_run_g2(cx.ws2, sh, bq.rest_tab2.p, w, 1, n, bq.rest.map.p, 1, st2, nullptr);
            G16_CUDA(cudaMemcpyAsync(out2, cx.ws2.result.p, (size_t)rows * sizeof(G2XYZZ), cudaMemcpyDeviceToDevice, st2));
        } else {
            G16_CUDA(cudaMemsetAsync(out2, 0, (size_t)rows * sizeof(G2XYZZ), st2));
        }
        cx.bitq_tmp2.ensure(rows);
        msm_sum_rows_g2(cx.ws2, bq.table2.p, cx.bitq_entries.p, rows * bq.groups, rows, cx.bitq_tmp2.p, st2);
        xyzz_add_g2(out2, cx.bitq_tmp2.p, rows, st2);
    } else {
        run_query_g1(cx.ws1b, st2, cx.qB, w, 1, n, true, rows, cx.resB1.p + sb, nullptr);
        if (sb + rows == n) G16_CUDA(cudaEventRecord(cx.ev_ab, st2));   // A and B1 of the whole batch are complete
        MsmShape sh = msm_make_shape(cx.nB2, rows, cx.cB2, 1);
        msm_run_g2(cx.ws2, sh, cx.tabB2.p, w, 1, n, cx.qB.map.p, 1, st2, nullptr);
        G16_CUDA(cudaMemcpyAsync(cx.resB2.p + sb, cx.ws2.result.p, (size_t)rows * sizeof(G2XYZZ), cudaMemcpyDeviceToDevice, st2));
    }
    if (bitq && cx.bqK.on) run_query_bitq_g1(cx, st2, cx.bqK, w, n, rows, true, cx.resK.p + sb);
    else run_query_g1(cx.ws1b, st2, cx.qK, w, 1, n, true, rows, cx.resK.p + sb, nullptr);
    if (cx.n_commit)   // proof of knowledge of the commitment: same scalars over BasisExpSigma
        run_query_g1(cx.ws1b, st2, cx.qPedSigma, w, 1, n, true, rows, cx.resPok.p + sb, nullptr);
}

// returns device ms (sum over stages). Leaves proofs in d_proofs. Throws on unsatisfied witness.
static float ctx_run_batch(Ctx& cx, size_t n, int kind) {
    if (!cx.solver_supported) throw std::runtime_error("unsupported circuit: " + cx.solver_unsupported_reason);
    cudaStream_t st = cx.stream;
    ctx_build_tables(cx);
    ctx_ensure_batch(cx, n);
    // Balanced sub-batches: a batch that needs k sub-batches is cut into k equal ones (multiples of 32) instead of full ones
    // and a remainder — 768 proofs run as 2 x 384 on the two lanes (96.5 ms) rather than 512 + 256 (99.5 ms on one stream).
    struct SubBatchGuard { Ctx& c; uint32_t saved; ~SubBatchGuard() { c.sub_batch = saved; } } sub_guard{cx, cx.sub_batch};
    if (n > cx.sub_batch) {
        const size_t k = (n + cx.sub_batch - 1) / cx.sub_batch;
        const size_t b = (((n + k - 1) / k) + 31) / 32 * 32;
        if (b < cx.sub_batch) cx.sub_batch = (uint32_t)b;
    }
    // combination tables of the wire-driven queries: built once the classification has seen enough witnesses
    if (cx.bitq_state == 0 && !cx.bitq_built && cx.bitq_rows_seen >= cx.bitq_min_rows && n >= cx.bitq_min_batch) {
        ctx_bitq_build(cx);
        if (cx.bitq_state == 0) cx.bitq_state = 1;
    }
    const bool bitq_profiling = cx.bitq_state == 0 && !cx.bitq_built && cx.bitq_rows_seen < cx.bitq_min_rows;
    const bool bitq_live = cx.bitq_state == 1 && cx.bitq_built && n >= cx.bitq_min_batch;
    if (bitq_live) G16_CUDA(cudaMemsetAsync(cx.d_bitq_exc.p, 0, sizeof(uint32_t), st));
    // evaluation-basis Z query (no commitment circuits only: their C evaluations are mostly full-width)
    // Not for the commitment (AES) circuits: their C evaluations are largely full-width, measured 1 112 vs 1 090 proofs/s.
    // Not when the caller asked for the coefficients of H (g16_prove_witness_detail).
    const bool eval_z = !cx.n_commit && !cx.want_h && (cx.eval_z > 0 || (cx.eval_z < 0 && n >= cx.eval_z_min));
    if (eval_z) ctx_build_eval_tables(cx);
    StageTimer& tm = cx.timer;
    tm.reset();
    size_t l0 = cx.ws1.launches + cx.ws1c.launches + cx.ws1b.launches + cx.ws2.launches + cx.dom.launches + cx.wsf[0].launches + cx.wsf[1].launches;
    size_t own = 0;
    cx.ws1.log_reset();
    cx.ws1c.log_reset();
    cx.ws1b.log_reset();
    cx.wsf[0].log_reset();
    cx.wsf[1].log_reset();
    cx.ws2.log_reset();
    const bool piped = cx.pipeline && n > cx.sub_batch;
    tm.mark(piped ? ST_COUNT : ST_SOLVE, st);   // pipelined: one interval that only counts towards the total
    G16_CUDA(cudaMemsetAsync(cx.d_status.p, 0, n * sizeof(uint32_t), st));
    G16_CUDA(cudaMemsetAsync(cx.Aev.p, 0, n * cx.n_dom * sizeof(Fr), st));
    G16_CUDA(cudaMemsetAsync(cx.Bev.p, 0, n * cx.n_dom * sizeof(Fr), st));
    G16_CUDA(cudaMemsetAsync(cx.Cev.p, 0, n * cx.n_dom * sizeof(Fr), st));
    if (kind == 1) {
        launch_chacha_witness(cx.d_keys.p, cx.d_nonces.p, cx.d_counters.p, cx.d_inputs.p, (uint32_t)n, cx.W.p, n, cx.d_ct.p, st);
    } else if (kind == 2) {
        launch_aes_witness(cx.d_keys.p, cx.staged_key_len, cx.d_nonces.p, cx.d_counters.p, cx.d_inputs.p, (uint32_t)n, cx.W.p, n,
                           cx.d_ct.p, st);
    } else {
        launch_witness_copy(cx.d_witness.p, cx.n_public - 1 + cx.n_secret, (uint32_t)n, cx.W.p, n, st);
    }
    launch_scalars_from_be(cx.d_rs_be.p, (uint32_t)(2 * n), cx.d_rs.p, st);
    own += 2;
    // W is wire-major: wire k of proof i at W[k*n + i]
    cudaStream_t st2 = cx.stream2;
    if (!piped) {
        // one main stream with stage timers: solve everything, then per sub-batch H and the Z query; the wire-driven
        // queries run on a side stream and fill the SM slots that the long H / Z kernels leave idle.
        // G16_SPLIT_SOLVE=1: only the first sub-batch is solved on the main stream; the later ones are solved on the side
        // stream while the main stream already runs the transforms and the Z query of the one before (the solver is a chain
        // of ~160 short level kernels: latency, not throughput). MEASURED ON B200 AND NOT ADOPTED: the solve interval
        // shrinks (6.85 -> 4.38 ms per 1024 proofs) but the side stream's wire queries start later and the same SM time is
        // taken out of the accumulate / reduce intervals instead: 145.6 vs 145.3 ms per step.
        // G16_SPLIT_SOLVE=2 (the default without a commitment): the sub-batches are solved CONCURRENTLY, the first on the main
        // stream and the others on the side stream, and the main stream goes on when all of them are done: two chains of short
        // level kernels fill each other's launch gaps and tails (solve 6.83 -> 6.26 ms per 1024 proofs).
        const bool split_solve = n > cx.sub_batch && cx.split_solve == 1;
        const bool par_solve = n > cx.sub_batch && cx.split_solve == 2;
        const uint32_t rows0 = (split_solve || par_solve) ? (uint32_t)cx.sub_batch : (uint32_t)n;
        if (par_solve && cx.solve_chains > 2 && !cx.n_commit) {
            // more than two chains (G16_SOLVE_CHAINS): the rows are cut into equal parts, one stream each
            const size_t P = cx.solve_chains;
            const size_t part = ((n + P - 1) / P + 31) / 32 * 32;
            while (cx.solve_streams.size() + 3 < P) { cudaStream_t s; G16_CUDA(cudaStreamCreate(&s)); cx.solve_streams.push_back(s); }
            G16_CUDA(cudaEventRecord(cx.ev_fork, st));   // witness assignment done
            size_t k = 0;
            for (size_t c = 1; c * part < n; c++, k++) {
                cudaStream_t s = c == 1 ? st2 : (c == 2 ? cx.stream3 : cx.solve_streams[c - 3]);
                const size_t sb = c * part;
                const uint32_t rows = (uint32_t)((n - sb) < part ? (n - sb) : part);
                G16_CUDA(cudaStreamWaitEvent(s, cx.ev_fork, 0));
                own += ctx_solve(cx, n, sb, rows, s, cx.ws1b);
                if (cx.ev_solved.size() <= k) { cudaEvent_t e; G16_CUDA(cudaEventCreate(&e)); cx.ev_solved.push_back(e); }
                G16_CUDA(cudaEventRecord(cx.ev_solved[k], s));
            }
            own += ctx_solve(cx, n, 0, (uint32_t)(part < n ? part : n), st, cx.ws1c);
            for (size_t i = 0; i < k; i++) G16_CUDA(cudaStreamWaitEvent(st, cx.ev_solved[i], 0));
        } else {
        if (par_solve) {
            G16_CUDA(cudaEventRecord(cx.ev_fork, st));   // witness assignment done
            G16_CUDA(cudaStreamWaitEvent(st2, cx.ev_fork, 0));
            size_t k = 0;
            for (size_t sb = cx.sub_batch; sb < n; sb += cx.sub_batch, k++) {
                uint32_t rows = (uint32_t)((n - sb) < cx.sub_batch ? (n - sb) : cx.sub_batch);
                own += ctx_solve(cx, n, sb, rows, st2, cx.ws1b);
            }
            if (cx.ev_solved.empty()) { cudaEvent_t e; G16_CUDA(cudaEventCreate(&e)); cx.ev_solved.push_back(e); }
            G16_CUDA(cudaEventRecord(cx.ev_solved[0], st2));
        }
        own += ctx_solve(cx, n, 0, rows0, st, par_solve ? cx.ws1c : cx.ws1b);
        if (par_solve) G16_CUDA(cudaStreamWaitEvent(st, cx.ev_solved[0], 0));
        }
        G16_CUDA(cudaEventRecord(cx.ev_fork, st));
        G16_CUDA(cudaStreamWaitEvent(st2, cx.ev_fork, 0));
        if (split_solve) {
            size_t k = 0;
            for (size_t sb = cx.sub_batch; sb < n; sb += cx.sub_batch, k++) {
                uint32_t rows = (uint32_t)((n - sb) < cx.sub_batch ? (n - sb) : cx.sub_batch);
                own += ctx_solve(cx, n, sb, rows, st2, cx.ws1b);
                if (cx.ev_solved.size() <= k) { cudaEvent_t e; G16_CUDA(cudaEventCreate(&e)); cx.ev_solved.push_back(e); }
                G16_CUDA(cudaEventRecord(cx.ev_solved[k], st2));
            }
        }
        if (bitq_profiling) {   // which wires held only 0 / 1 in this batch (read back below, folded into cx.bit_mask)
            cx.d_bit_flags.ensure(cx.nb_wires);
            G16_CUDA(cudaMemsetAsync(cx.d_bit_flags.p, 3, cx.nb_wires * sizeof(uint32_t), st2));   // every byte 3: the low two bits are what counts
            bitq_profile(cx.W.p, n, cx.nb_wires, (uint32_t)n, cx.d_bit_flags.p, st2);
        }
        for (size_t sb = 0; sb < n; sb += cx.sub_batch) {
            uint32_t rows = (uint32_t)((n - sb) < cx.sub_batch ? (n - sb) : cx.sub_batch);
            Fr* a = cx.Aev.p + sb * cx.n_dom;
            if (split_solve && sb > 0) G16_CUDA(cudaStreamWaitEvent(st, cx.ev_solved[sb / cx.sub_batch - 1], 0));
            tm.mark(ST_H, st);
            if (eval_z) {
                compute_d_run(cx.dom, a, cx.Bev.p + sb * cx.n_dom, cx.n_dom, rows, st);
                z_query_ba_policy(cx.ws1, cx.qQd, rows);
                run_query_g1(cx.ws1, st, cx.qQd, a, cx.n_dom, 1, false, rows, cx.resZ.p + sb, &tm);
            } else {
                compute_h_run(cx.dom, a, cx.Bev.p + sb * cx.n_dom, cx.Cev.p + sb * cx.n_dom, cx.n_dom, rows, st);
                // h (gnark order) pairs index-for-index with G1.Z; the other queries read wire values through their maps
                z_query_ba_policy(cx.ws1, cx.qZ, rows);
                run_query_g1(cx.ws1, st, cx.qZ, a, cx.n_dom, 1, false, rows, cx.resZ.p + sb, &tm);
            }
            ctx_wire_queries(cx, n, sb, rows, st2, eval_z);
        }
    } else {
        // Pipelined schedule: sub-batch k runs its whole chain (solve -> H -> Z query) on lane k % 2, so the latency-bound
        // phases of one sub-batch (solver levels, bucket sort, reduction trees) overlap the integer-multiply-bound phases
        // (NTT, bucket accumulation) of the other lane; the wire-driven queries follow on the side stream as soon as the
        // sub-batch's witness is complete. Stage timers are meaningless here: only the total is measured.
        if (cx.test_lane_oom) {   // G16_TEST_LANE_OOM=1: what an allocation failure of the second lane's scratch looks like (tests of the fallback)
            cx.test_lane_oom = false;
            throw CudaError("cudaMalloc: out of memory (G16_TEST_LANE_OOM)");
        }
        cudaStream_t lane[2] = {st, cx.stream3};
        MsmWorkspace<G1>* wsz[2] = {&cx.ws1, &cx.ws1c};
        G16_CUDA(cudaEventRecord(cx.ev_fork, st));
        G16_CUDA(cudaStreamWaitEvent(lane[1], cx.ev_fork, 0));
        size_t k = 0;
        for (size_t sb = 0; sb < n; sb += cx.sub_batch, k++) {
            uint32_t rows = (uint32_t)((n - sb) < cx.sub_batch ? (n - sb) : cx.sub_batch);
            cudaStream_t s = lane[k & 1];
            MsmWorkspace<G1>& wz = *wsz[k & 1];
            own += ctx_solve(cx, n, sb, rows, s, wz);
            if (cx.ev_solved.size() <= k) { cudaEvent_t e; G16_CUDA(cudaEventCreate(&e)); cx.ev_solved.push_back(e); }
            G16_CUDA(cudaEventRecord(cx.ev_solved[k], s));
            G16_CUDA(cudaStreamWaitEvent(st2, cx.ev_solved[k], 0));
            ctx_wire_queries(cx, n, sb, rows, st2, eval_z);
            Fr* a = cx.Aev.p + sb * cx.n_dom;
            // stagger the lanes: the transforms of this sub-batch start when those of the one before are done, so that they
            // run beside its digit sort (L2 atomics, no multiplier) and this sub-batch's sort beside its bucket additions
            if (k > 0 && cx.pipeline_stagger) G16_CUDA(cudaStreamWaitEvent(s, cx.ev_hdone[k - 1], 0));
            if (eval_z) compute_d_run(cx.dom, a, cx.Bev.p + sb * cx.n_dom, cx.n_dom, rows, s);
            else compute_h_run(cx.dom, a, cx.Bev.p + sb * cx.n_dom, cx.Cev.p + sb * cx.n_dom, cx.n_dom, rows, s);
            if (cx.ev_hdone.size() <= k) { cudaEvent_t e; G16_CUDA(cudaEventCreate(&e)); cx.ev_hdone.push_back(e); }
            G16_CUDA(cudaEventRecord(cx.ev_hdone[k], s));
            z_query_ba_policy(wz, eval_z ? cx.qQd : cx.qZ, rows);
            run_query_g1(wz, s, eval_z ? cx.qQd : cx.qZ, a, cx.n_dom, 1, false, rows, cx.resZ.p + sb, nullptr);
        }
        if (bitq_profiling) {   // the side stream has waited for every sub-batch's witness by now
            cx.d_bit_flags.ensure(cx.nb_wires);
            G16_CUDA(cudaMemsetAsync(cx.d_bit_flags.p, 3, cx.nb_wires * sizeof(uint32_t), st2));
            bitq_profile(cx.W.p, n, cx.nb_wires, (uint32_t)n, cx.d_bit_flags.p, st2);
        }
        G16_CUDA(cudaEventRecord(cx.ev_join3, lane[1]));
        G16_CUDA(cudaStreamWaitEvent(st, cx.ev_join3, 0));
    }
    // Ar, Bs1, -rs delta and the half-products s*Ar, r*Bs1 only need the A and B1 results: they run on their own stream beside
    // the K / B2 queries and the Z query (G16_ASSEMBLE_EARLY=0: after the last MSM, on the main stream)
    static const int asm_early = env_int("G16_ASSEMBLE_EARLY", 1);
    const bool early = asm_early && assemble_team_enabled();
    if (early) {
        G16_CUDA(cudaStreamWaitEvent(cx.stream4, cx.ev_ab, 0));
        own += launch_assemble_products(cx.keys, cx.asm_scratch, (uint32_t)n, cx.resA.p, cx.resB1.p, cx.d_rs.p, cx.stream4);
        G16_CUDA(cudaEventRecord(cx.ev_prod, cx.stream4));
    }
    // the G2 element only needs the G2 MSM: assembled on the side stream, off the critical path of the G1 chain
    launch_assemble_g2(cx.keys, (uint32_t)n, cx.resB2.p, cx.d_rs.p, cx.d_proofs.p, cx.proof_bytes(), st2);
    own += 1;
    G16_CUDA(cudaEventRecord(cx.ev_join, st2));
    G16_CUDA(cudaStreamWaitEvent(st, cx.ev_join, 0));
    tm.mark(ST_ASSEMBLE, st);
    if (eval_z) { xyzz_add_g1(cx.resZ.p, cx.resZc.p, (uint32_t)n, st); own += 1; }
    if (early) {
        G16_CUDA(cudaStreamWaitEvent(st, cx.ev_prod, 0));
        own += launch_assemble_finish(cx.asm_scratch, cx.n_commit != 0, (uint32_t)n, cx.resK.p, cx.resZ.p, cx.d_proofs.p, cx.proof_bytes(), st);
    } else {
        own += launch_assemble(cx.keys, cx.asm_scratch, cx.n_commit != 0, (uint32_t)n, cx.resA.p, cx.resB1.p, cx.resK.p, cx.resZ.p,
                               cx.d_rs.p, cx.d_proofs.p, cx.proof_bytes(), st);
    }
    if (cx.n_commit) {
        launch_assemble_commitment(cx.commit_aff.p, cx.resPok.p, (uint32_t)n, cx.d_proofs.p, cx.proof_bytes(), st);
        own += 1;
    }
    tm.mark(-1, st);
    cx.h_status.assign(n, 0);
    cx.d_status.download(cx.h_status.data(), n, st);
    G16_CUDA(cudaStreamSynchronize(st));
    uint32_t status = 0;
    for (uint32_t v : cx.h_status) status |= v;
    float per[ST_COUNT];
    float total = tm.finish(per);
    for (int i = 0; i < ST_COUNT; i++) cx.stage_ms[i] = per[i];
    cx.stage_ms[6] = total;
    cx.launches = own + (cx.ws1.launches + cx.ws1c.launches + cx.ws1b.launches + cx.ws2.launches + cx.dom.launches + cx.wsf[0].launches + cx.wsf[1].launches - l0);
    cx.stage_ms[7] = (float)cx.launches;
    G16_CUDA(cudaStreamSynchronize(st2));
    G16_CUDA(cudaStreamSynchronize(cx.stream3));
    G16_CUDA(cudaStreamSynchronize(cx.stream4));
    if (bitq_profiling && !(status & 7u)) {
        std::vector<uint32_t> f(cx.nb_wires);
        cx.d_bit_flags.download(f.data(), cx.nb_wires, st);
        G16_CUDA(cudaStreamSynchronize(st));
        if (cx.bit_mask.empty()) cx.bit_mask.assign(cx.nb_wires, 3u);
        for (uint32_t w = 0; w < cx.nb_wires; w++) cx.bit_mask[w] &= f[w] & 3u;
        cx.bitq_rows_seen += n;
    }
    if (bitq_live) {
        uint32_t exc = 0;
        cx.d_bitq_exc.download(&exc, 1, st);
        G16_CUDA(cudaStreamSynchronize(st));
        if (cx.bitq_test_exception) { exc = 1; cx.bitq_test_exception = false; }
        if (exc) {
            // a wire classified as a bit (trit) held something else: the table sums of this batch are wrong. Prove it again on
            // the general path. The classification was learned, not proved, so learn on: the mask keeps what it knows (it is
            // only ever narrowed), the next bitq_min_rows witnesses narrow it further and the tables are rebuilt from it. After
            // three such rounds the tables are switched off for good.
            cx.bqA = BitQuery(); cx.bqB = BitQuery(); cx.bqK = BitQuery();
            cx.bitq_built = false;
            cx.bitq_rows_seen = 0;
            cx.bitq_state = ++cx.bitq_relearned > 3 ? 2 : 0;
            return ctx_run_batch(cx, n, kind);
        }
    }
    cx.counters[12] = (uint64_t)cx.bqA.rest.n | ((uint64_t)cx.bqB.rest.n << 20) | ((uint64_t)cx.bqK.rest.n << 40);   // points left on the general path
    cx.counters[11] = (uint64_t)cx.bitq_state | ((uint64_t)(bitq_live ? 1 : 0) << 8) | ((uint64_t)cx.bqA.groups << 16) | ((uint64_t)cx.bqK.groups << 40);
    cx.counters[6] = cx.ws1.log_sum(st) + cx.ws1c.log_sum(cx.stream3);   // G1 mixed additions of the Z query (lanes)
    cx.counters[0] = cx.counters[6] + cx.ws1b.log_sum(st2) + cx.wsf[0].log_sum(cx.fan[0]) + cx.wsf[1].log_sum(cx.fan[2]);               // ... of all G1 queries
    cx.counters[1] = cx.ws2.log_sum(st2);   // G2 mixed additions
    cx.counters[2] = cx.ws1.log_n + cx.ws1c.log_n + cx.ws1b.log_n;   // G1 accumulate launches
    cx.counters[3] = cx.ws2.log_n;
    cx.counters[4] = cx.launches;
    cx.counters[5] = n;
    cx.counters[8] = cx.ws1.log_sum(st, 1) + cx.ws1c.log_sum(cx.stream3, 1);   // sorted slots of the Z query (entries + batch-affine padding)
    cx.counters[13] = cx.ws1.log_sum(st, 2) + cx.ws1c.log_sum(cx.stream3, 2);   // entries of the XYZZ accumulation of the Z query (group sums + direct leftovers)
    cx.counters[9] = (uint64_t)cx.ws1.last_K;                                   // batch-affine levels of the Z query
    cx.counters[10] = (uint64_t)cx.nZ_buckets_total(n);
    cx.counters[7] = (uint64_t)cx.sub_batch | ((uint64_t)(piped ? 1 : 0) << 32) | ((uint64_t)(eval_z ? 1 : 0) << 33);
    if (status & 4u) throw std::runtime_error("solver: unsupported hint");
    // Per-witness verdicts stay in h_status: the batch has run to completion, so the proofs of the satisfied witnesses are
    // valid and can still be fetched (g16_chacha_batch_fetch zeroes the others); the call itself reports G16_ERR_UNSAT.
    if (status & 3u) {
        size_t bad = 0, first = n;
        for (size_t i = 0; i < n; i++) if (cx.h_status[i] & 3u) { if (first == n) first = i; bad++; }
        throw std::domain_error("witness does not satisfy the constraint system (" + std::to_string(bad) + " of " + std::to_string(n) +
                                " requests, first at index " + std::to_string(first) + ")");
    }
    return total;
}

}  // namespace g16
