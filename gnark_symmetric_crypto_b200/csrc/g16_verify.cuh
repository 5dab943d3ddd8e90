// Verifier context: one verifying key resident on one GPU, and batched groth16.Verify. Host code only.
// Replaces gnark v0.11.0 backend/groth16/bn254/verify.go (Verify) and marshal.go (VerifyingKey.ReadFrom), as driven by
// libraries/verifier/impl/verify_impl.go:36-58 (key loading) and verifiers.go:50-152 (ChachaVerifier / AESVerifier) —
// SURVEY.md §8f rank 4, Appendix B (vk bytes), C (proof bytes), F.4 (equation).
//
//   kSum_i = K_0 + sum_k pub_{i,k} K_k (+ challenge_i K_last + C_i with a BSB22 commitment)    one batched MSM (msm.cuh)
//   accept <=> e(-Ar, Bs) e(alpha, beta2) e(kSum, gamma2) e(Krs, delta2) == 1                    pairing.cuh
//              (and, with a commitment, e(C, G) e(PoK, GRootSigmaNeg) == 1)
// Every proof gets its own verdict (the reference's Verify answers per proof), so no random linear combination is used.
#pragma once
#include "common.cuh"
#include "host_parse.hpp"
#include "msm_types.hpp"
#include "ntt_api.hpp"
#include "pairing_api.hpp"
#include "prover_api.hpp"
#include <memory>

namespace g16 {

struct VCtx {
    int device = 0;
    cudaStream_t stream = nullptr;
    uint32_t nK = 0, n_public = 0, n_commit = 0;   // n_public: public inputs without ONE and without the commitment challenge
    VerifyKeys keys;
    DevBuf<G1Affine> K, tabK;
    int cK = 13;
    MsmWorkspace<G1> ws;
    PairingWorkspace pw;
    // batch state
    DevBuf<uint8_t> d_proofs, d_pub_be, ok1, ok2, verdict;
    DevBuf<Fr> d_pub, W;
    DevBuf<G1Affine> P, P2, commit, commit_tmp;
    DevBuf<G2Affine> Q, Q2;
    DevBuf<G1XYZZ> commit_x;
    DevBuf<uint32_t> bad;
    float last_ms = 0.f;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~VCtx() {
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (stream) cudaStreamDestroy(stream);
    }
    size_t proof_bytes() const { return n_commit ? 196 : 164; }
};

static std::unique_ptr<VCtx> vctx_create(const uint8_t* vk_bytes, size_t vk_len, int device) {
    std::unique_ptr<VCtx> v(new VCtx());
    v->device = device;
    G16_CUDA(cudaSetDevice(device));
    G16_CUDA(cudaStreamCreate(&v->stream));
    G16_CUDA(cudaEventCreate(&v->e0));
    G16_CUDA(cudaEventCreate(&v->e1));
    cudaStream_t st = v->stream;
    VkFile vk = parse_vk(vk_bytes, vk_len);
    if (vk.ped.size() > 1) throw ParseError("vk: only circuits with at most one BSB22 commitment are supported");
    for (auto& l : vk.public_and_commitment_committed)
        if (!l.empty()) throw ParseError("vk: public-committed wires are not supported");
    v->n_commit = (uint32_t)vk.ped.size();
    v->nK = vk.nK;
    if (vk.nK < 1 + v->n_commit) throw ParseError("vk: len(G1.K) too small");
    v->n_public = vk.nK - 1 - v->n_commit;
    DevBuf<uint32_t> err(1);
    err.zero(st);
    {
        DevBuf<uint8_t> raw;
        raw.upload(vk.K, (size_t)vk.nK * 32, st);
        v->K.alloc(vk.nK);
        launch_decompress_g1(raw.p, vk.nK, v->K.p, err.p, st);
        G16_CUDA(cudaStreamSynchronize(st));
    }
    // the six key points (and the Pedersen verification key) through the same kernels
    std::vector<uint8_t> g1raw(32), g2raw(64 * 5, 0);
    memcpy(g1raw.data(), vk.alpha, 32);
    memcpy(&g2raw[0], vk.beta2, 64);
    memcpy(&g2raw[64], vk.gamma2, 64);
    memcpy(&g2raw[128], vk.delta2, 64);
    g2raw[192] = g2raw[256] = 0x40;   // infinity unless a commitment key follows
    if (v->n_commit) { memcpy(&g2raw[192], vk.ped[0].g, 64); memcpy(&g2raw[256], vk.ped[0].g_root_sigma_neg, 64); }
    DevBuf<uint8_t> d1, d2;
    DevBuf<G1Affine> a1(1);
    DevBuf<G2Affine> a2(5);
    d1.upload(g1raw.data(), 32, st);
    d2.upload(g2raw.data(), 320, st);
    launch_decompress_g1(d1.p, 1, a1.p, err.p, st);
    launch_decompress_g2(d2.p, 5, a2.p, err.p, st);
    uint32_t herr = 0;
    G2Affine h2[5];
    a1.download(&v->keys.alpha, 1, st);
    a2.download(h2, 5, st);
    err.download(&herr, 1, st);
    G16_CUDA(cudaStreamSynchronize(st));
    if (herr) throw ParseError("vk: point decompression failed (flags=" + std::to_string(herr) + ")");
    {   // VerifyingKey.ReadFrom checks the subgroup of its G2 points as well
        pairing_consts_ensure(v->pw, st);
        DevBuf<uint8_t> okb(5);
        launch_g2_subgroup(a2.p, 5, v->pw.consts.p, okb.p, st);
        uint8_t hok[5];
        okb.download(hok, 5, st);
        G16_CUDA(cudaStreamSynchronize(st));
        for (uint8_t x : hok)
            if (!x) throw ParseError("vk: a G2 point is not in the correct subgroup");
    }
    v->keys.beta2 = h2[0]; v->keys.gamma2 = h2[1]; v->keys.delta2 = h2[2]; v->keys.ped_g = h2[3]; v->keys.ped_gneg = h2[4];
    v->keys.n_commit = v->n_commit;
    // fixed-base table of K: one bucket set per proof, like the prover's wire-driven queries
    const int nwin = (254 + v->cK - 1) / v->cK;
    v->tabK.alloc((size_t)vk.nK * nwin);
    msm_precompute_g1(v->K.p, vk.nK, nwin, v->cK, v->tabK.p, st);
    G16_CUDA(cudaStreamSynchronize(st));
    return v;
}

// proofs: n x proof_bytes (host). pub: n x n_public field elements (host), Montgomery 4 x u64 (fmt 0) or 32-byte big-endian
// canonical (fmt 1). ok_out: n bytes, 1 = accepted. Returns device ms.
static float vctx_verify_batch(VCtx& v, size_t n, const uint8_t* proofs, const void* pub, int fmt, uint8_t* ok_out) {
    cudaStream_t st = v.stream;
    G16_CUDA(cudaSetDevice(v.device));
    const size_t pb = v.proof_bytes();
    const uint32_t np = v.n_public, nw = np + v.n_commit;   // scalars per proof besides ONE
    const size_t SUB = 1024;                                  // bounds the MSM and line-record scratch (~70 MB per 1024 proofs)
    v.verdict.ensure(n);
    G16_CUDA(cudaEventRecord(v.e0, st));
    for (size_t sb = 0; sb < n; sb += SUB) {
        const uint32_t rows = (uint32_t)((n - sb) < SUB ? (n - sb) : SUB);
        v.d_proofs.upload(proofs + sb * pb, (size_t)rows * pb, st);
        v.d_pub.ensure((size_t)rows * (nw ? nw : 1));
        if (fmt == 1) {
            if (np) {
                v.d_pub_be.upload((const uint8_t*)pub + sb * np * 32, (size_t)rows * np * 32, st);
                fr_be_to_mont(v.d_pub_be.p, rows * np, v.d_pub.p, st);
            }
        } else if (np) {
            G16_CUDA(cudaMemcpyAsync(v.d_pub.p, (const uint8_t*)pub + sb * np * 32, (size_t)rows * np * 32, cudaMemcpyHostToDevice, st));
        }
        // W wire-major: W[k * rows + i] ; k = 0 is ONE, 1..np the public inputs, np + 1 the commitment challenge
        v.W.ensure((size_t)rows * v.nK);
        launch_witness_copy(v.d_pub.p, np, rows, v.W.p, rows, st);
        v.P.ensure((size_t)rows * 4); v.Q.ensure((size_t)rows * 4);
        v.bad.ensure(rows);
        G16_CUDA(cudaMemsetAsync(v.bad.p, 0, (size_t)rows * 4, st));
        if (v.n_commit) { v.P2.ensure((size_t)rows * 2); v.Q2.ensure((size_t)rows * 2); v.commit.ensure(rows); v.commit_tmp.ensure(rows); v.commit_x.ensure(rows); }
        pairing_consts_ensure(v.pw, st);
        launch_verify_unpack(v.keys, v.d_proofs.p, pb, rows, v.P.p, v.Q.p, v.P2.p, v.Q2.p, v.commit.p, v.pw.consts.p, v.bad.p, st);
        if (v.n_commit) {
            // challenge = hash_to_field(C) goes to the last scalar (prove.go does the same on the prover side, a11)
            launch_g1_affine_to_xyzz(v.commit.p, rows, v.commit_x.p, st);
            launch_bsb22_challenge(v.commit_x.p, rows, v.W.p, rows, np + 1, v.commit_tmp.p, st);
        }
        MsmShape sh = msm_make_shape(v.nK, rows, v.cK, 1);
        msm_run_g1(v.ws, sh, v.tabK.p, v.W.p, 1, rows, nullptr, 1, st, nullptr);
        launch_verify_ksum(v.ws.result.p, v.n_commit ? v.commit.p : nullptr, rows, v.P.p, st);
        v.ok1.ensure(rows);
        pairing_check_run(v.pw, v.P.p, v.Q.p, 4, rows, v.ok1.p, st);
        if (v.n_commit) {
            v.ok2.ensure(rows);
            pairing_check_run(v.pw, v.P2.p, v.Q2.p, 2, rows, v.ok2.p, st);
        }
        launch_verify_verdict(v.ok1.p, v.n_commit ? v.ok2.p : nullptr, v.bad.p, rows, v.verdict.p + sb, st);
        G16_CUDA(cudaStreamSynchronize(st));   // the staging buffers are reused by the next sub-batch
    }
    G16_CUDA(cudaEventRecord(v.e1, st));
    v.verdict.download(ok_out, n, st);
    G16_CUDA(cudaStreamSynchronize(st));
    G16_CUDA(cudaEventElapsedTime(&v.last_ms, v.e0, v.e1));
    return v.last_ms;
}

}  // namespace g16
