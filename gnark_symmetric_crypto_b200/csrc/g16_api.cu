// C-ABI of libg16b200.so (include/g16b200.h). Host code only. Every entry point converts exceptions into status codes;
// nothing propagates across the ABI. There is no CPU fallback: without a CUDA device these functions return G16_ERR_CUDA.
#include "../../include/g16b200.h"
#include "g16_ctx.cuh"
#include "g16_verify.cuh"
#include "pairing_api.hpp"
#include <mutex>
#include <random>
#include <thread>

using namespace g16;

#if !defined(G16_EMU)
// CUDA 12 loads kernels lazily: the first launch of every kernel variant (tree arities, merge levels, ... chosen by the batch
// size) stalls for its load, 100-300 ms for the large hot kernels. A serving process meets new batch sizes for a long time
// (measured: 1 s stalls in the middle of a 1024-caller run). Ask for eager loading unless the host process decided
// otherwise; this runs when the library is loaded, before its first CUDA call.
__attribute__((constructor)) static void g16_eager_module_loading() { setenv("CUDA_MODULE_LOADING", "EAGER", 0); }
#endif

// One handle = the context of one GPU (cx) plus, for a multi-device handle (g16_init_multi), the handles of the other
// devices of the list. Independent proofs are the sharding unit: request i of a batch goes to device i mod G,
// every device runs its own pipeline from its own host thread, results are gathered in input order; no collective.
// `mu` guards one entry point at a time on the device context; `seq_mu` is held across the stage -> run -> fetch sequence of
// the one-call batch entry points, so that callers sharing a device (the multi-device handle used by ProveBatch, the per-device
// handles used by the Prove batcher's workers) never interleave their phases. (The split-phase API used by benchmarks is
// sequenced by its caller.)
struct DevCore {
    std::unique_ptr<Ctx> cx;
    std::mutex mu, seq_mu;
};
struct g16_ctx {
    std::shared_ptr<DevCore> core = std::make_shared<DevCore>();
    std::unique_ptr<Ctx>& cx = core->cx;
    std::mutex& mu = core->mu;
    std::vector<std::unique_ptr<g16_ctx>> extra;   // devices[1..] of a multi-device handle
    std::unique_ptr<g16_ctx> view0;                // single-device handle of devices[0] (shares this handle's core)
    std::vector<size_t> shard_n;                   // requests staged on each device by the last multi-device stage
    size_t staged_total = 0;
    g16_ctx() {}
    explicit g16_ctx(std::shared_ptr<DevCore> c) : core(std::move(c)) {}
    size_t ndev() const { return 1 + extra.size(); }
    // single-device handle of slot k: for slot 0 of a multi-device handle a view that shares the core but never shards
    g16_ctx* dev(size_t k) {
        if (k > 0) return extra[k - 1].get();
        if (extra.empty()) return this;
        if (!view0) view0.reset(new g16_ctx(core));
        return view0.get();
    }
};
// locks the phase sequence of every device a batch of n requests will touch (slot order: no deadlock between callers)
struct SeqLock {
    std::vector<std::unique_lock<std::mutex>> held;
    SeqLock(g16_ctx* ctx, size_t n) {
        if (!ctx) return;
        const size_t G = (ctx->ndev() > 1 && n >= 2) ? ctx->ndev() : 1;
        for (size_t k = 0; k < G; k++) held.emplace_back(ctx->dev(k)->core->seq_mu);
    }
};

struct g16_vctx {
    std::unique_ptr<VCtx> v;
    std::mutex mu;
};

struct g16_msm_plan {
    int group, device, c;
    uint32_t n;
    DevBuf<G1Affine> p1, t1;   // points ; window tables 2^(c w) P_i (precomputed mode)
    DevBuf<G2Affine> p2, t2;
    int precomp = 0;
    DevBuf<Fr> scalars;
    int scalars_mont = 0;
    MsmWorkspace<G1> ws1;
    MsmWorkspace<G2> ws2;
    DevBuf<G1Affine> o1;
    DevBuf<G2Affine> o2;
    StageTimer tm;
    cudaStream_t stream = nullptr;
    ~g16_msm_plan() { if (stream) cudaStreamDestroy(stream); }
};

static thread_local std::string t_last_error;

template <class F>
static int guarded(F&& fn) {
    try {
        fn();
        t_last_error.clear();
        return G16_OK;
    } catch (const ParseError& e) {
        t_last_error = e.what();
        return G16_ERR_PARSE;
    } catch (const CudaError& e) {
        t_last_error = e.what();
        return G16_ERR_CUDA;
    } catch (const std::domain_error& e) {
        t_last_error = e.what();
        return G16_ERR_UNSAT;
    } catch (const std::invalid_argument& e) {
        t_last_error = e.what();
        return G16_ERR_ARG;
    } catch (const std::bad_alloc&) {
        t_last_error = "out of host memory";
        return G16_ERR_ARG;
    } catch (const std::exception& e) {
        t_last_error = e.what();
        return G16_ERR_UNSUPPORTED;
    } catch (...) {
        t_last_error = "unknown error";
        return G16_ERR_UNSUPPORTED;
    }
}
#define REQUIRE(cond, msg) do { if (!(cond)) throw std::invalid_argument(msg); } while (0)

static void require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) throw CudaError("no CUDA device available (libg16b200 has no CPU fallback)");
}

// OS CSPRNG, rejection-sampled into [0, r)  (gnark draws r, s with fr.Element.SetRandom from crypto/rand)
static void random_scalars_be(uint8_t* out, size_t count) {
    static const uint8_t R_BE[32] = {0x30, 0x64, 0x4e, 0x72, 0xe1, 0x31, 0xa0, 0x29, 0xb8, 0x50, 0x45, 0xb6, 0x81, 0x81, 0x58, 0x5d,
                                     0x28, 0x33, 0xe8, 0x48, 0x79, 0xb9, 0x70, 0x91, 0x43, 0xe1, 0xf5, 0x93, 0xf0, 0x00, 0x00, 0x01};
    std::random_device rd("/dev/urandom");
    for (size_t i = 0; i < count; i++) {
        uint8_t* o = out + 32 * i;
        for (;;) {
            for (int k = 0; k < 32; k += 4) {
                uint32_t v = rd();
                memcpy(o + k, &v, 4);
            }
            o[0] &= 0x3F;
            if (memcmp(o, R_BE, 32) < 0) break;
        }
    }
}

// rs: per proof r | s (2 x 32 B big-endian) followed, for circuits with a hints.Randomize wire (AES), by the 32-byte
// commitment mask: 64 or 96 bytes per proof. NULL = draw everything from the OS CSPRNG.
static size_t rs_bytes_per_proof(const Ctx& c) { return c.has_randomize ? 96 : 64; }
static void stage_masks(Ctx& c, size_t n, const uint8_t* masks_be) {
    c.d_mask_be.upload(masks_be, 32 * n, c.stream);
    c.d_mask.ensure(n);
    fr_be_to_mont(c.d_mask_be.p, (uint32_t)n, c.d_mask.p, c.stream);
}
static void stage_rs(Ctx& c, size_t n, const uint8_t* rs) {
    const size_t per = rs_bytes_per_proof(c);
    std::vector<uint8_t> tmp, rs64(64 * n), masks;
    if (!rs) {
        tmp.resize(per * n);
        random_scalars_be(tmp.data(), per / 32 * n);
        rs = tmp.data();
    }
    for (size_t i = 0; i < n; i++) memcpy(&rs64[64 * i], rs + per * i, 64);
    c.d_rs_be.upload(rs64.data(), 64 * n, c.stream);
    if (c.has_randomize) {
        masks.resize(32 * n);
        for (size_t i = 0; i < n; i++) memcpy(&masks[32 * i], rs + per * i + 64, 32);
        stage_masks(c, n, masks.data());
    }
    G16_CUDA(cudaStreamSynchronize(c.stream));   // host temporaries must outlive the copies
    // r, s and the commitment masks are the proof's zero-knowledge randomness: do not leave copies on the host heap
    auto wipe = [](std::vector<uint8_t>& v) { if (!v.empty()) { volatile uint8_t* p = v.data(); for (size_t i = 0; i < v.size(); i++) p[i] = 0; } };
    wipe(tmp); wipe(rs64); wipe(masks);
}

static void prove_witness_impl(g16_ctx* ctx, const uint64_t* witness, size_t n_witness, const uint8_t* rs,
                               uint8_t* proof_out, size_t* proof_len, uint64_t* msm_g1_out, uint64_t* msm_g2_out,
                               uint64_t* h_out) {
    REQUIRE(ctx && witness && proof_out, "NULL argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    Ctx& c = *ctx->cx;
    REQUIRE(n_witness == (size_t)c.n_public - 1 + c.n_secret, "witness length must be nbPublic-1+nbSecret");
    G16_CUDA(cudaSetDevice(c.device));
    c.d_witness.upload((const Fr*)witness, n_witness, c.stream);
    stage_rs(c, 1, rs);
    c.staged = 1;
    c.staged_kind = 0;
    struct WantH { Ctx& c; ~WantH() { c.want_h = false; } } want_h_guard{c};
    c.want_h = h_out != nullptr;
    ctx_run_batch(c, 1, 0);
    c.d_proofs.download(proof_out, c.proof_bytes(), c.stream);
    if (proof_len) *proof_len = c.proof_bytes();
    if (msm_g1_out) {
        DevBuf<G1Affine> aff(4);
        DevBuf<G1XYZZ> x(4);
        const G1XYZZ* src[4] = {c.resA.p, c.resB1.p, c.resK.p, c.resZ.p};
        for (int i = 0; i < 4; i++) G16_CUDA(cudaMemcpyAsync(x.p + i, src[i], sizeof(G1XYZZ), cudaMemcpyDeviceToDevice, c.stream));
        xyzz_to_affine_g1(x.p, 4, aff.p, c.stream);
        aff.download((G1Affine*)msm_g1_out, 4, c.stream);
        G16_CUDA(cudaStreamSynchronize(c.stream));
    }
    if (msm_g2_out) {
        DevBuf<G2Affine> aff(1);
        xyzz_to_affine_g2(c.resB2.p, 1, aff.p, c.stream);
        aff.download((G2Affine*)msm_g2_out, 1, c.stream);
        G16_CUDA(cudaStreamSynchronize(c.stream));
    }
    if (h_out) c.Aev.download((Fr*)h_out, c.n_dom, c.stream);
    G16_CUDA(cudaStreamSynchronize(c.stream));
}

static int log2_exact(size_t n) {
    int k = 0;
    while (((size_t)1 << k) < n) k++;
    if (((size_t)1 << k) != n) throw std::invalid_argument("n must be a power of two");
    return k;
}

extern "C" {

int g16_version(void) { return 100; }
const char* g16_last_error(void) { return t_last_error.c_str(); }
int g16_device_count(int* n) {
    return guarded([&] {
        REQUIRE(n, "n is NULL");
        G16_CUDA(cudaGetDeviceCount(n));
    });
}

int g16_init(const uint8_t* pk, size_t pk_len, const uint8_t* r1cs, size_t r1cs_len, int device, g16_ctx** out) {
    return guarded([&] {
        REQUIRE(pk && r1cs && out, "NULL argument");
        require_device();
        std::unique_ptr<g16_ctx> h(new g16_ctx());
        h->cx = ctx_create(pk, pk_len, r1cs, r1cs_len, device);
        *out = h.release();
    });
}
// Same key on every device of the list (a device may appear more than once: two contexts on one GPU, which is how the
// sharding is tested on a one-GPU box). The contexts are created concurrently, one host thread per device.
int g16_init_multi(const uint8_t* pk, size_t pk_len, const uint8_t* r1cs, size_t r1cs_len, const int* devices, size_t n_devices,
                   g16_ctx** out) {
    return guarded([&] {
        REQUIRE(pk && r1cs && out && devices, "NULL argument");
        REQUIRE(n_devices >= 1 && n_devices <= 64, "device list must hold 1..64 entries");
        require_device();
        int have = 0;
        G16_CUDA(cudaGetDeviceCount(&have));
        for (size_t k = 0; k < n_devices; k++) REQUIRE(devices[k] >= 0 && devices[k] < have, "device index out of range");
        std::vector<std::unique_ptr<g16_ctx>> hs(n_devices);
        std::vector<std::exception_ptr> errs(n_devices);
        std::vector<std::thread> th;
        for (size_t k = 0; k < n_devices; k++)
            th.emplace_back([&, k] {
                try {
                    hs[k].reset(new g16_ctx());
                    hs[k]->cx = ctx_create(pk, pk_len, r1cs, r1cs_len, devices[k]);
                } catch (...) {
                    errs[k] = std::current_exception();
                }
            });
        for (auto& t : th) t.join();
        auto release_all = [&] {
            for (auto& h : hs)
                if (h && h->cx) { cudaSetDevice(h->cx->device); h.reset(); }
        };
        for (size_t k = 0; k < n_devices; k++)
            if (errs[k]) { release_all(); std::rethrow_exception(errs[k]); }
        std::unique_ptr<g16_ctx> head = std::move(hs[0]);
        for (size_t k = 1; k < n_devices; k++) head->extra.push_back(std::move(hs[k]));
        head->dev(0);   // create the slot-0 view now (not lazily from concurrent callers)
        *out = head.release();
    });
}
int g16_ctx_devices(const g16_ctx* ctx, int* devices_out, size_t cap, size_t* n_out) {
    return guarded([&] {
        REQUIRE(ctx && n_out, "NULL argument");
        g16_ctx* c = const_cast<g16_ctx*>(ctx);
        *n_out = c->ndev();
        for (size_t k = 0; devices_out && k < c->ndev() && k < cap; k++) devices_out[k] = c->dev(k)->cx->device;
    });
}
int g16_ctx_device_handle(g16_ctx* ctx, size_t k, g16_ctx** out) {
    return guarded([&] {
        REQUIRE(ctx && out, "NULL argument");
        REQUIRE(k < ctx->ndev(), "device slot out of range");
        *out = ctx->dev(k);
    });
}
void g16_free(g16_ctx* ctx) {
    if (!ctx) return;
    try {
        for (auto& e : ctx->extra) {
            cudaSetDevice(e->cx->device);
            e.reset();
        }
        cudaSetDevice(ctx->cx->device);
        delete ctx;
    } catch (...) {
    }
}
int g16_info(const g16_ctx* ctx, uint64_t info[16]) {
    return guarded([&] {
        REQUIRE(ctx && info, "NULL argument");
        std::lock_guard<std::mutex> lk(const_cast<g16_ctx*>(ctx)->mu);
        const Ctx& c = *ctx->cx;
        uint64_t v[16] = {c.n_dom, c.nA, c.nB, c.nZ, c.nK, c.nB2, c.nb_wires, c.n_public, c.n_secret, c.n_constraints,
                          c.n_instr, c.nlevels, c.n_commit, c.proof_bytes(), (uint64_t)c.device,
                          (uint64_t)(c.solver_supported ? 1 : 0)};
        memcpy(info, v, sizeof(v));
    });
}

}  // extern "C"

// ---- single-device phases (the multi-device entry points below run them per device)
static int chacha_stage_one(g16_ctx* ctx, size_t n, const uint8_t* keys, const uint8_t* nonces, const uint32_t* counters,
                            const uint8_t* inputs, const uint8_t* rs) {
    return guarded([&] {
        REQUIRE(ctx && keys && nonces && counters && inputs, "NULL argument");
        REQUIRE(n > 0 && n <= (1u << 20), "batch size out of range");
        std::lock_guard<std::mutex> lk(ctx->mu);
        Ctx& c = *ctx->cx;
        REQUIRE(c.n_public == 1153 && c.n_secret == 256, "context does not hold the ChaCha20 circuit");
        G16_CUDA(cudaSetDevice(c.device));
        c.d_keys.upload(keys, 32 * n, c.stream);
        c.d_nonces.upload(nonces, 12 * n, c.stream);
        c.d_counters.upload(counters, n, c.stream);
        c.d_inputs.upload(inputs, 64 * n, c.stream);
        stage_rs(c, n, rs);
        c.staged = n;
        c.staged_kind = 1;
    });
}
static int batch_run_one(g16_ctx* ctx, float* ms) {
    return guarded([&] {
        REQUIRE(ctx, "NULL argument");
        std::lock_guard<std::mutex> lk(ctx->mu);
        Ctx& c = *ctx->cx;
        REQUIRE(c.staged > 0 && c.staged_kind != 0, "no staged request batch");
        G16_CUDA(cudaSetDevice(c.device));
        float t;
        try {
            t = ctx_run_batch(c, c.staged, c.staged_kind);
        } catch (const CudaError& e) {
            // the second lane's scratch (another ~28 GB for a 512-proof sub-batch) did not fit next to whatever else lives on this
            // GPU: give it back and prove the batch on the single main stream from now on
            if (!c.pipeline || !strstr(e.what(), "out of memory")) throw;
            cudaGetLastError();
            cudaDeviceSynchronize();
            c.ws1c = MsmWorkspace<G1>();
            c.pipeline = 0;
            t = ctx_run_batch(c, c.staged, c.staged_kind);
        }
        if (ms) *ms = t;
    });
}
static int batch_fetch_one(g16_ctx* ctx, uint8_t* proofs_out, uint8_t* ct_out) {
    return guarded([&] {
        REQUIRE(ctx && proofs_out, "NULL argument");
        std::lock_guard<std::mutex> lk(ctx->mu);
        Ctx& c = *ctx->cx;
        REQUIRE(c.staged > 0, "no staged batch");
        G16_CUDA(cudaSetDevice(c.device));
        const size_t pb = c.proof_bytes();
        c.d_proofs.download(proofs_out, c.staged * pb, c.stream);
        if (ct_out) c.d_ct.download(ct_out, c.staged * 64, c.stream);
        G16_CUDA(cudaStreamSynchronize(c.stream));
        // a request whose witness did not satisfy the system has no proof: its slot is zeroed (g16_last_batch_status)
        for (size_t i = 0; i < c.staged && i < c.h_status.size(); i++)
            if (c.h_status[i]) memset(proofs_out + i * pb, 0, pb);
    });
}
static int aes_stage_one(g16_ctx* ctx, size_t n, const uint8_t* keys, size_t key_len, const uint8_t* nonces,
                         const uint32_t* counters, const uint8_t* inputs, const uint8_t* rsm) {
    return guarded([&] {
        REQUIRE(ctx && keys && nonces && counters && inputs, "NULL argument");
        REQUIRE(n > 0 && n <= (1u << 20), "batch size out of range");
        REQUIRE(key_len == 16 || key_len == 32, "key length must be 16 or 32");
        std::lock_guard<std::mutex> lk(ctx->mu);
        Ctx& c = *ctx->cx;
        REQUIRE(c.n_public == 142 && c.n_secret == key_len, "context does not hold the AES circuit for this key length");
        G16_CUDA(cudaSetDevice(c.device));
        c.d_keys.upload(keys, key_len * n, c.stream);
        c.d_nonces.upload(nonces, 12 * n, c.stream);
        c.d_counters.upload(counters, n, c.stream);
        c.d_inputs.upload(inputs, 64 * n, c.stream);
        stage_rs(c, n, rsm);
        c.staged = n;
        c.staged_kind = 2;
        c.staged_key_len = (uint32_t)key_len;
    });
}

// ---- multi-device plumbing: request i of the batch -> device i mod G (slot i / G on that device)
static size_t shard_count(size_t n, size_t G, size_t k) { return n / G + (k < n % G ? 1 : 0); }
template <class T>
static std::vector<T> shard_gather(const T* src, size_t n, size_t per, size_t G, size_t k) {
    std::vector<T> out(shard_count(n, G, k) * per);
    size_t j = 0;
    for (size_t i = k; i < n; i += G, j++) memcpy(&out[j * per], src + i * per, per * sizeof(T));
    return out;
}
// kind 1 = chacha (key_len 32), 2 = aes
static int multi_stage(g16_ctx* ctx, int kind, size_t n, const uint8_t* keys, size_t key_len, const uint8_t* nonces,
                       const uint32_t* counters, const uint8_t* inputs, const uint8_t* rs) {
    if (!ctx) { t_last_error = "NULL argument"; return G16_ERR_ARG; }
    const size_t G = ctx->ndev();
    if (G == 1 || n < 2) {
        ctx->shard_n.clear();
        ctx->staged_total = n;
        return kind == 1 ? chacha_stage_one(ctx, n, keys, nonces, counters, inputs, rs)
                         : aes_stage_one(ctx, n, keys, key_len, nonces, counters, inputs, rs);
    }
    if (!keys || !nonces || !counters || !inputs) { t_last_error = "NULL argument"; return G16_ERR_ARG; }
    const size_t rs_per = rs_bytes_per_proof(*ctx->cx);
    std::vector<size_t> shards(G, 0);
    for (size_t k = 0; k < G; k++) {
        const size_t m = shard_count(n, G, k);
        shards[k] = m;
        if (!m) continue;
        auto k_ = shard_gather(keys, n, key_len, G, k);
        auto n_ = shard_gather(nonces, n, 12, G, k);
        auto c_ = shard_gather(counters, n, 1, G, k);
        auto i_ = shard_gather(inputs, n, 64, G, k);
        std::vector<uint8_t> r_;
        if (rs) r_ = shard_gather(rs, n, rs_per, G, k);
        int rc = kind == 1 ? chacha_stage_one(ctx->dev(k), m, k_.data(), n_.data(), c_.data(), i_.data(), rs ? r_.data() : nullptr)
                           : aes_stage_one(ctx->dev(k), m, k_.data(), key_len, n_.data(), c_.data(), i_.data(), rs ? r_.data() : nullptr);
        if (rc) return rc;
    }
    ctx->shard_n = shards;
    ctx->staged_total = n;
    return G16_OK;
}
static int multi_run(g16_ctx* ctx, float* ms) {
    if (!ctx) { t_last_error = "NULL argument"; return G16_ERR_ARG; }
    if (ctx->shard_n.empty()) return batch_run_one(ctx, ms);
    const size_t G = ctx->ndev();
    std::vector<int> rcs(G, 0);
    std::vector<float> t(G, 0.f);
    std::vector<std::string> msg(G);
    std::vector<std::thread> th;
    for (size_t k = 0; k < G; k++) {
        if (!ctx->shard_n[k]) continue;
        th.emplace_back([&, k] {
            rcs[k] = batch_run_one(ctx->dev(k), &t[k]);
            if (rcs[k]) msg[k] = t_last_error;   // thread-local in the worker: carried back by hand
        });
    }
    for (auto& x : th) x.join();
    float mx = 0.f;
    int rc = 0;
    for (size_t k = 0; k < G; k++) {
        if (t[k] > mx) mx = t[k];
        // an unsatisfied witness on one device does not stop the others; any other failure wins
        if (rcs[k] && (rc == 0 || rc == G16_ERR_UNSAT)) { rc = rcs[k]; t_last_error = msg[k]; }
    }
    if (ms) *ms = mx;   // the devices run concurrently: the batch takes as long as the slowest
    return rc;
}
static int multi_fetch(g16_ctx* ctx, uint8_t* proofs_out, uint8_t* ct_out) {
    if (!ctx) { t_last_error = "NULL argument"; return G16_ERR_ARG; }
    if (ctx->shard_n.empty()) return batch_fetch_one(ctx, proofs_out, ct_out);
    if (!proofs_out) { t_last_error = "NULL argument"; return G16_ERR_ARG; }
    const size_t G = ctx->ndev(), pb = ctx->cx->proof_bytes();
    for (size_t k = 0; k < G; k++) {
        const size_t m = ctx->shard_n[k];
        if (!m) continue;
        std::vector<uint8_t> p(m * pb), c(m * 64);
        int rc = batch_fetch_one(ctx->dev(k), p.data(), c.data());
        if (rc) return rc;
        for (size_t j = 0; j < m; j++) {
            memcpy(proofs_out + (k + j * G) * pb, &p[j * pb], pb);
            if (ct_out) memcpy(ct_out + (k + j * G) * 64, &c[j * 64], 64);
        }
    }
    return G16_OK;
}

extern "C" {

int g16_chacha_batch_stage(g16_ctx* ctx, size_t n, const uint8_t* keys, const uint8_t* nonces, const uint32_t* counters,
                           const uint8_t* inputs, const uint8_t* rs) {
    return multi_stage(ctx, 1, n, keys, 32, nonces, counters, inputs, rs);
}
int g16_aes_batch_stage(g16_ctx* ctx, size_t n, const uint8_t* keys, size_t key_len, const uint8_t* nonces,
                        const uint32_t* counters, const uint8_t* inputs, const uint8_t* rsm) {
    return multi_stage(ctx, 2, n, keys, key_len, nonces, counters, inputs, rsm);
}
int g16_chacha_batch_run(g16_ctx* ctx, float* ms) { return multi_run(ctx, ms); }
int g16_chacha_batch_fetch(g16_ctx* ctx, uint8_t* proofs_out, uint8_t* ct_out) { return multi_fetch(ctx, proofs_out, ct_out); }
// A batch with unsatisfiable requests still proves the others: the call returns G16_ERR_UNSAT after writing every proof
// that exists (the slots of the failed requests are zero); g16_last_batch_status tells which ones failed.
int g16_prove_chacha_batch(g16_ctx* ctx, size_t n, const uint8_t* keys, const uint8_t* nonces, const uint32_t* counters,
                           const uint8_t* inputs, const uint8_t* rs, uint8_t* proofs_out, uint8_t* ct_out) {
    SeqLock seq(ctx, n);
    int rc = g16_chacha_batch_stage(ctx, n, keys, nonces, counters, inputs, rs);
    if (rc) return rc;
    rc = g16_chacha_batch_run(ctx, nullptr);
    if (rc && rc != G16_ERR_UNSAT) return rc;
    std::string keep = t_last_error;
    int rc2 = g16_chacha_batch_fetch(ctx, proofs_out, ct_out);
    if (rc2) return rc2;
    t_last_error = keep;
    return rc;
}
int g16_prove_aes_batch(g16_ctx* ctx, size_t n, const uint8_t* keys, size_t key_len, const uint8_t* nonces,
                        const uint32_t* counters, const uint8_t* inputs, const uint8_t* rsm, uint8_t* proofs_out, uint8_t* ct_out) {
    SeqLock seq(ctx, n);
    int rc = g16_aes_batch_stage(ctx, n, keys, key_len, nonces, counters, inputs, rsm);
    if (rc) return rc;
    rc = g16_chacha_batch_run(ctx, nullptr);
    if (rc && rc != G16_ERR_UNSAT) return rc;
    std::string keep = t_last_error;
    int rc2 = g16_chacha_batch_fetch(ctx, proofs_out, ct_out);
    if (rc2) return rc2;
    t_last_error = keep;
    return rc;
}
// status_out[i] for request i of the last batch run on this context: 0 = proved; bit 0 an unsatisfied constraint, bit 1 a
// division by zero while solving (both: G16_ERR_UNSAT for that request), bit 2 an unsupported instruction.
int g16_last_batch_status(g16_ctx* ctx, uint32_t* status_out, size_t n) {
    return guarded([&] {
        REQUIRE(ctx && status_out, "NULL argument");
        if (ctx->shard_n.empty()) {
            std::lock_guard<std::mutex> lk(ctx->mu);
            REQUIRE(n <= ctx->cx->h_status.size(), "n exceeds the size of the last batch");
            memcpy(status_out, ctx->cx->h_status.data(), n * sizeof(uint32_t));
            return;
        }
        const size_t G = ctx->ndev();
        REQUIRE(n <= ctx->staged_total, "n exceeds the size of the last batch");
        for (size_t i = 0; i < n; i++) {
            g16_ctx* d = ctx->dev(i % G);
            std::lock_guard<std::mutex> lk(d->mu);
            status_out[i] = i / G < d->cx->h_status.size() ? d->cx->h_status[i / G] : 0u;
        }
    });
}
int g16_set_schedule(g16_ctx* ctx, int pipeline, int sub_batch) {
    return guarded([&] {
        REQUIRE(ctx, "NULL argument");
        REQUIRE(sub_batch >= 0 && sub_batch <= (1 << 16), "sub-batch out of range");
        for (size_t k = 0; k < ctx->ndev(); k++) {
            std::lock_guard<std::mutex> lk(ctx->dev(k)->mu);
            Ctx& c = *ctx->dev(k)->cx;
            c.pipeline = pipeline ? 1 : 0;
            if (sub_batch > 0) c.sub_batch = (uint32_t)sub_batch;
        }
    });
}
int g16_last_stage_ms(const g16_ctx* ctx, float ms[8]) {
    return guarded([&] {
        REQUIRE(ctx && ms, "NULL argument");
        std::lock_guard<std::mutex> lk(const_cast<g16_ctx*>(ctx)->mu);
        memcpy(ms, ctx->cx->stage_ms, sizeof(float) * 8);
    });
}

int g16_last_counters(const g16_ctx* ctx, uint64_t out[8]) {
    return guarded([&] {
        REQUIRE(ctx && out, "NULL argument");
        g16_ctx* h = const_cast<g16_ctx*>(ctx);
        {
            std::lock_guard<std::mutex> lk(h->mu);
            memcpy(out, ctx->cx->counters, sizeof(uint64_t) * 8);
        }
        if (!h->shard_n.empty())   // multi-device batch: work counters summed over the devices that took part
            for (size_t k = 1; k < h->ndev(); k++) {
                if (!h->shard_n[k]) continue;
                std::lock_guard<std::mutex> lk(h->dev(k)->mu);
                for (int j = 0; j < 7; j++) out[j] += h->dev(k)->cx->counters[j];
            }
    });
}

int g16_last_counters_ex(const g16_ctx* ctx, uint64_t out[16]) {
    return guarded([&] {
        REQUIRE(ctx && out, "NULL argument");
        g16_ctx* h = const_cast<g16_ctx*>(ctx);
        std::lock_guard<std::mutex> lk(h->mu);
        memcpy(out, ctx->cx->counters, sizeof(uint64_t) * 16);
    });
}

int g16_prove_witness(g16_ctx* ctx, const uint64_t* witness, size_t n_witness, const uint8_t* rs, uint8_t* proof_out,
                      size_t* proof_len) {
    return guarded([&] { prove_witness_impl(ctx, witness, n_witness, rs, proof_out, proof_len, nullptr, nullptr, nullptr); });
}
int g16_prove_witness_detail(g16_ctx* ctx, const uint64_t* witness, size_t n_witness, const uint8_t* rs, uint8_t* proof_out,
                             size_t* proof_len, uint64_t* msm_g1_out, uint64_t* msm_g2_out, uint64_t* h_out) {
    return guarded([&] { prove_witness_impl(ctx, witness, n_witness, rs, proof_out, proof_len, msm_g1_out, msm_g2_out, h_out); });
}

int g16_solve_ex(g16_ctx* ctx, const uint64_t* witness, size_t n_witness, size_t batch, const uint8_t* masks_be, uint64_t* W,
                 uint64_t* A, uint64_t* B, uint64_t* C) {
    return guarded([&] {
        REQUIRE(ctx && witness && batch > 0, "bad argument");
        std::lock_guard<std::mutex> lk(ctx->mu);
        Ctx& c = *ctx->cx;
        REQUIRE(n_witness == (size_t)c.n_public - 1 + c.n_secret, "witness length must be nbPublic-1+nbSecret");
        if (!c.solver_supported) throw std::runtime_error("unsupported circuit: " + c.solver_unsupported_reason);
        REQUIRE(masks_be || !c.has_randomize, "this circuit has a hints.Randomize wire: masks are required");
        G16_CUDA(cudaSetDevice(c.device));
        cudaStream_t st = c.stream;
        if (c.n_commit) ctx_build_tables(c);
        ctx_ensure_batch(c, batch);
        c.d_witness.upload((const Fr*)witness, n_witness * batch, st);
        if (c.has_randomize) stage_masks(c, batch, masks_be);
        G16_CUDA(cudaMemsetAsync(c.d_status.p, 0, batch * sizeof(uint32_t), st));
        G16_CUDA(cudaMemsetAsync(c.Aev.p, 0, batch * c.n_dom * sizeof(Fr), st));
        G16_CUDA(cudaMemsetAsync(c.Bev.p, 0, batch * c.n_dom * sizeof(Fr), st));
        G16_CUDA(cudaMemsetAsync(c.Cev.p, 0, batch * c.n_dom * sizeof(Fr), st));
        launch_witness_copy(c.d_witness.p, (uint32_t)n_witness, (uint32_t)batch, c.W.p, batch, st);
        ctx_solve(c, batch, 0, (uint32_t)batch, st, c.ws1b);
        c.h_status.assign(batch, 0);
        c.d_status.download(c.h_status.data(), batch, st);
        DevBuf<Fr> rows;
        if (W) {
            rows.alloc(batch * c.nb_wires);
            launch_wires_to_rows(c.W.p, batch, (uint32_t)batch, c.nb_wires, rows.p, st);
            rows.download((Fr*)W, batch * c.nb_wires, st);
        }
        for (size_t i = 0; i < batch; i++) {
            if (A) G16_CUDA(cudaMemcpyAsync(A + i * 4 * c.n_constraints, c.Aev.p + i * c.n_dom, (size_t)c.n_constraints * 32, cudaMemcpyDeviceToHost, st));
            if (B) G16_CUDA(cudaMemcpyAsync(B + i * 4 * c.n_constraints, c.Bev.p + i * c.n_dom, (size_t)c.n_constraints * 32, cudaMemcpyDeviceToHost, st));
            if (C) G16_CUDA(cudaMemcpyAsync(C + i * 4 * c.n_constraints, c.Cev.p + i * c.n_dom, (size_t)c.n_constraints * 32, cudaMemcpyDeviceToHost, st));
        }
        G16_CUDA(cudaStreamSynchronize(st));
        uint32_t status = 0;
        for (uint32_t v : c.h_status) status |= v;
        if (status & 4u) throw std::runtime_error("solver: unsupported hint");
        if (status & 3u) throw std::domain_error("witness does not satisfy the constraint system");
    });
}
int g16_solve(g16_ctx* ctx, const uint64_t* witness, size_t n_witness, size_t batch, uint64_t* W, uint64_t* A, uint64_t* B,
              uint64_t* C) {
    return g16_solve_ex(ctx, witness, n_witness, batch, nullptr, W, A, B, C);
}

int g16_aes_witness(const uint8_t* keys, size_t key_len, const uint8_t* nonces, const uint32_t* counters, const uint8_t* inputs,
                    size_t n, uint8_t* ct_out, uint64_t* witness_out) {
    return guarded([&] {
        REQUIRE(keys && nonces && counters && inputs && n > 0 && n <= (1u << 20), "bad argument");
        REQUIRE(key_len == 16 || key_len == 32, "key length must be 16 or 32");
        require_device();
        const uint32_t nw = 142 + (uint32_t)key_len;   // ONE | Nonce[12] | Counter | Plaintext[64] | Ciphertext[64] | Key
        DevBuf<uint8_t> dk, dn, di, dct(64 * n);
        DevBuf<uint32_t> dc;
        DevBuf<Fr> W((size_t)nw * n), rows((size_t)nw * n);
        dk.upload(keys, key_len * n); dn.upload(nonces, 12 * n); dc.upload(counters, n); di.upload(inputs, 64 * n);
        launch_aes_witness(dk.p, (uint32_t)key_len, dn.p, dc.p, di.p, (uint32_t)n, W.p, n, dct.p, 0);
        if (ct_out) dct.download(ct_out, 64 * n);
        if (witness_out) {
            launch_wires_to_rows(W.p, n, (uint32_t)n, nw, rows.p, 0);
            rows.download((Fr*)witness_out, (size_t)nw * n);
        }
        G16_CUDA(cudaDeviceSynchronize());
    });
}
int g16_bsb22_challenge(const uint64_t* commitments, size_t n, uint64_t* challenges_out) {
    return guarded([&] {
        REQUIRE(commitments && challenges_out && n > 0 && n <= (1u << 20), "bad argument");
        require_device();
        DevBuf<G1Affine> aff, aff2(n);
        DevBuf<G1XYZZ> xy(n);
        DevBuf<Fr> out(n);
        aff.upload((const G1Affine*)commitments, n);
        launch_g1_affine_to_xyzz(aff.p, (uint32_t)n, xy.p, 0);
        launch_bsb22_challenge(xy.p, (uint32_t)n, out.p, n, 0, aff2.p, 0);   // "wire 0" of an n-wide witness = out[i]
        out.download((Fr*)challenges_out, n);
        G16_CUDA(cudaDeviceSynchronize());
    });
}

int g16_compute_h(g16_ctx* ctx, const uint64_t* a, const uint64_t* b, const uint64_t* c_in, uint64_t* h_out) {
    return guarded([&] {
        REQUIRE(ctx && a && b && c_in && h_out, "NULL argument");
        std::lock_guard<std::mutex> lk(ctx->mu);
        Ctx& c = *ctx->cx;
        G16_CUDA(cudaSetDevice(c.device));
        cudaStream_t st = c.stream;
        ctx_ensure_batch(c, 1);
        const uint64_t* src[3] = {a, b, c_in};
        Fr* dst[3] = {c.Aev.p, c.Bev.p, c.Cev.p};
        for (int i = 0; i < 3; i++) {
            G16_CUDA(cudaMemsetAsync(dst[i], 0, c.n_dom * sizeof(Fr), st));
            G16_CUDA(cudaMemcpyAsync(dst[i], src[i], (size_t)c.n_constraints * 32, cudaMemcpyHostToDevice, st));
        }
        compute_h_run(c.dom, c.Aev.p, c.Bev.p, c.Cev.p, c.n_dom, 1, st);
        c.Aev.download((Fr*)h_out, c.n_dom, st);
        G16_CUDA(cudaStreamSynchronize(st));
    });
}

// ------------------------------------------------------------------------------------------------ stage-level: fields, groups
int g16_field_op(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    return guarded([&] {
        REQUIRE(a && out && (field == 0 || field == 1) && op >= 0 && op <= 7, "bad argument");
        REQUIRE(b || op >= 3, "binary op needs b");
        require_device();
        DevBuf<uint64_t> da, db, dout(4 * n);
        da.upload(a, 4 * n);
        if (b) db.upload(b, 4 * n);
        launch_field_op(field, op, da.p, b ? db.p : nullptr, dout.p, n, 0);
        dout.download(out, 4 * n);
        G16_CUDA(cudaDeviceSynchronize());
    });
}
int g16_group_op(int group, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n) {
    return guarded([&] {
        REQUIRE(a && out && (group == 1 || group == 2) && op >= 0 && op <= (group == 1 ? 5 : 3), "bad argument");
        REQUIRE(b || op == 2, "op needs b");
        require_device();
        size_t pw = group == 1 ? 8 : 16;
        size_t bw = (op == 0 || op >= 3) ? pw : 4;
        DevBuf<uint64_t> da, db, dout(pw * n);
        da.upload(a, pw * n);
        if (b) db.upload(b, bw * n);
        launch_group_op(group, op, da.p, b ? db.p : nullptr, dout.p, n, 0);
        dout.download(out, pw * n);
        G16_CUDA(cudaDeviceSynchronize());
    });
}
int g16_decompress(int group, const uint8_t* in, uint64_t* out, size_t n) {
    return guarded([&] {
        REQUIRE(in && out && (group == 1 || group == 2), "bad argument");
        require_device();
        DevBuf<uint8_t> din;
        DevBuf<uint32_t> err(1);
        err.zero();
        din.upload(in, n * (group == 1 ? 32 : 64));
        if (group == 1) {
            DevBuf<G1Affine> o(n);
            launch_decompress_g1(din.p, (uint32_t)n, o.p, err.p, 0);
            o.download((G1Affine*)out, n);
            G16_CUDA(cudaDeviceSynchronize());
        } else {
            DevBuf<G2Affine> o(n);
            launch_decompress_g2(din.p, (uint32_t)n, o.p, err.p, 0);
            o.download((G2Affine*)out, n);
            G16_CUDA(cudaDeviceSynchronize());
        }
        uint32_t e = 0;
        err.download(&e, 1);
        G16_CUDA(cudaDeviceSynchronize());
        if (e) throw ParseError("point decompression failed (flags=" + std::to_string(e) + ")");
    });
}

// ------------------------------------------------------------------------------------------------ stage-level: MSM
int g16_msm_plan_create(int group, const uint64_t* points, size_t n, int window, int device, g16_msm_plan** out) {
    return guarded([&] {
        REQUIRE(points && out && (group == 1 || group == 2), "bad argument");
        REQUIRE(n > 0 && n <= (1u << 26), "n out of range");
        require_device();
        G16_CUDA(cudaSetDevice(device));
        std::unique_ptr<g16_msm_plan> p(new g16_msm_plan());
        p->group = group; p->device = device; p->n = (uint32_t)n;
        p->c = window > 0 ? window : msm_pick_window(n);
        REQUIRE(p->c >= 2 && p->c <= 24, "window out of range");
        G16_CUDA(cudaStreamCreate(&p->stream));
        if (group == 1) { p->p1.upload((const G1Affine*)points, n, p->stream); p->o1.alloc(1); }
        else { p->p2.upload((const G2Affine*)points, n, p->stream); p->o2.alloc(1); }
        p->scalars.alloc(n);
        G16_CUDA(cudaStreamSynchronize(p->stream));
        *out = p.release();
    });
}
// Fixed bases (an SRS): tabulate 2^(c w) P_i once, so that every later MSM needs one bucket set instead of one per window
// and no doubling chain at the end — the mode the prover context uses for the pk queries.
int g16_msm_plan_precompute(g16_msm_plan* plan, int window) {
    return guarded([&] {
        REQUIRE(plan, "NULL argument");
        G16_CUDA(cudaSetDevice(plan->device));
        int c = window;
        if (c <= 0) {   // argmin over c of ceil(254/c) * n + 2^c (accumulation + one bucket reduction)
            double best = 1e300;
            for (int k = 4; k <= 24; k++) {
                double cost = (double)((254 + k - 1) / k) * (double)plan->n + (double)(1u << k);
                if (cost < best) { best = cost; c = k; }
            }
        }
        REQUIRE(c >= 2 && c <= 24, "window out of range");
        const int nwin = (254 + c - 1) / c;
        REQUIRE((size_t)plan->n * nwin < (1ull << 31), "table too large for 31-bit point references");
        if (plan->group == 1) {
            plan->t1.alloc((size_t)plan->n * nwin);
            msm_precompute_g1(plan->p1.p, plan->n, nwin, c, plan->t1.p, plan->stream);
        } else {
            plan->t2.alloc((size_t)plan->n * nwin);
            msm_precompute_g2(plan->p2.p, plan->n, nwin, c, plan->t2.p, plan->stream);
        }
        G16_CUDA(cudaStreamSynchronize(plan->stream));
        plan->c = c;
        plan->precomp = 1;
    });
}
int g16_msm_plan_set_scalars(g16_msm_plan* plan, const uint64_t* scalars, int scalars_mont) {
    return guarded([&] {
        REQUIRE(plan && scalars, "NULL argument");
        G16_CUDA(cudaSetDevice(plan->device));
        plan->scalars.upload((const Fr*)scalars, plan->n, plan->stream);
        plan->scalars_mont = scalars_mont;
        G16_CUDA(cudaStreamSynchronize(plan->stream));
    });
}
int g16_msm_plan_run(g16_msm_plan* plan, uint64_t* out, float ms[4]) {
    return guarded([&] {
        REQUIRE(plan && out, "NULL argument");
        G16_CUDA(cudaSetDevice(plan->device));
        cudaStream_t st = plan->stream;
        MsmShape sh = msm_make_shape(plan->n, 1, plan->c, plan->precomp);
        plan->tm.reset();
        if (plan->group == 1) {
            msm_run_g1(plan->ws1, sh, plan->precomp ? plan->t1.p : plan->p1.p, plan->scalars.p, plan->n, 1, nullptr, plan->scalars_mont, st,
                       &plan->tm);
            xyzz_to_affine_g1(plan->ws1.result.p, 1, plan->o1.p, st);
            plan->o1.download((G1Affine*)out, 1, st);
        } else {
            msm_run_g2(plan->ws2, sh, plan->precomp ? plan->t2.p : plan->p2.p, plan->scalars.p, plan->n, 1, nullptr, plan->scalars_mont, st,
                       &plan->tm);
            xyzz_to_affine_g2(plan->ws2.result.p, 1, plan->o2.p, st);
            plan->o2.download((G2Affine*)out, 1, st);
        }
        G16_CUDA(cudaStreamSynchronize(st));
        if (ms) {
            float per[ST_COUNT];
            ms[0] = plan->tm.finish(per);
            ms[1] = per[ST_MSM_ACC];
            ms[2] = per[ST_MSM_SORT];
            ms[3] = per[ST_MSM_REDUCE];
        }
    });
}
void g16_msm_plan_free(g16_msm_plan* plan) {
    if (!plan) return;
    try {
        cudaSetDevice(plan->device);
        delete plan;
    } catch (...) {
    }
}
int g16_msm(int group, const uint64_t* points, const uint64_t* scalars, int scalars_mont, size_t n, int window, uint64_t* out,
            float ms[4]) {
    g16_msm_plan* p = nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    int rc = g16_msm_plan_create(group, points, n, window, dev, &p);
    if (rc) return rc;
    rc = g16_msm_plan_set_scalars(p, scalars, scalars_mont);
    if (!rc) rc = g16_msm_plan_run(p, out, ms);
    std::string keep = t_last_error;
    g16_msm_plan_free(p);
    t_last_error = keep;
    return rc;
}

// Stage-level view of the combination-table path (k_bitq.cu) for the parity tests: per-witness subset sums of `n` points.
// wires: [n][rows] Montgomery Fr, wire-major like the prover's witness array (wire i pairs with point i). The first n_binary
// wires must hold 0 or 1 (binary groups of 8 consecutive points, 255 subset sums each), the others 0, 1 or -1 (ternary groups of
// 5, 242 signed combinations); anything else raises *exception_out (the sums are then meaningless, as in the prover, which
// falls back). out: `rows` affine points, one table point per group and witness.
int g16_bitq_sum(int group, const uint64_t* points, size_t n, size_t n_binary, const uint64_t* wires, size_t rows, uint64_t* out,
                 uint32_t* exception_out) {
    return guarded([&] {
        REQUIRE(points && wires && out, "NULL argument");
        REQUIRE(group == 1 || group == 2, "group must be 1 (G1) or 2 (G2)");
        REQUIRE(n > 0 && n <= (1u << 22) && rows > 0 && rows <= (1u << 20) && n_binary <= n, "size out of range");
        require_device();
        cudaStream_t st = nullptr;
        const uint32_t groups_bin = (uint32_t)((n_binary + BITQ_K - 1) / BITQ_K);
        const uint32_t groups = groups_bin + (uint32_t)((n - n_binary + BITQ_T - 1) / BITQ_T);
        std::vector<uint32_t> ident((size_t)groups * BITQ_K, BITQ_NONE);
        for (size_t i = 0; i < n_binary; i++) ident[i] = (uint32_t)i;
        for (size_t i = n_binary; i < n; i++) ident[((size_t)groups_bin + (i - n_binary) / BITQ_T) * BITQ_K + (i - n_binary) % BITQ_T] = (uint32_t)i;
        DevBuf<uint32_t> d_ident, d_exc(1);
        DevBuf<Fr> d_w;
        DevBuf<uint2> d_entries((size_t)rows * groups);
        d_ident.upload(ident.data(), ident.size(), st);
        d_w.upload(reinterpret_cast<const Fr*>(wires), n * rows, st);
        d_exc.zero(st);
        bitq_entries(d_w.p, rows, (uint32_t)rows, d_ident.p, groups, groups_bin, d_entries.p, d_exc.p, st);
        if (group == 1) {
            DevBuf<G1Affine> pts, table((size_t)groups << BITQ_K), aff(rows);
            DevBuf<G1XYZZ> sums(rows);
            MsmWorkspace<G1> ws;
            pts.upload(reinterpret_cast<const G1Affine*>(points), n, st);
            bitq_build_g1(pts.p, d_ident.p, groups, groups_bin, table.p, st);
            msm_sum_rows_g1(ws, table.p, d_entries.p, (uint32_t)(rows * groups), (uint32_t)rows, sums.p, st);
            xyzz_to_affine_g1(sums.p, (uint32_t)rows, aff.p, st);
            aff.download(reinterpret_cast<G1Affine*>(out), rows, st);
            G16_CUDA(cudaStreamSynchronize(st));
        } else {
            DevBuf<G2Affine> pts, table((size_t)groups << BITQ_K), aff(rows);
            DevBuf<G2XYZZ> sums(rows);
            MsmWorkspace<G2> ws;
            pts.upload(reinterpret_cast<const G2Affine*>(points), n, st);
            bitq_build_g2(pts.p, d_ident.p, groups, groups_bin, table.p, st);
            msm_sum_rows_g2(ws, table.p, d_entries.p, (uint32_t)(rows * groups), (uint32_t)rows, sums.p, st);
            xyzz_to_affine_g2(sums.p, (uint32_t)rows, aff.p, st);
            aff.download(reinterpret_cast<G2Affine*>(out), rows, st);
            G16_CUDA(cudaStreamSynchronize(st));
        }
        uint32_t exc = 0;
        d_exc.download(&exc, 1, st);
        G16_CUDA(cudaStreamSynchronize(st));
        if (exception_out) *exception_out = exc;
    });
}

// ------------------------------------------------------------------------------------------------ setup
int g16_setup(const uint8_t* r1cs, size_t r1cs_len, const uint8_t* trapdoor_be, int device, uint8_t** pk_out, size_t* pk_len,
              uint8_t** vk_out, size_t* vk_len) {
    return guarded([&] {
        REQUIRE(r1cs && r1cs_len && pk_out && pk_len && vk_out && vk_len, "NULL argument");
        *pk_out = *vk_out = nullptr;
        *pk_len = *vk_len = 0;
        require_device();
        G16_CUDA(cudaSetDevice(device));
        uint8_t td[192];
        if (trapdoor_be) memcpy(td, trapdoor_be, 192);
        else random_scalars_be(td, 6);   // gnark draws the toxic waste from crypto/rand (zero has probability 2^-254)
        std::vector<uint8_t> pk, vk;
        cudaStream_t st = nullptr;
        G16_CUDA(cudaStreamCreate(&st));
        try {
            setup_run(r1cs, r1cs_len, td, pk, vk, st);
        } catch (...) {
            cudaStreamDestroy(st);
            volatile uint8_t* p = td; for (int i = 0; i < 192; i++) p[i] = 0;
            throw;
        }
        cudaStreamDestroy(st);
        { volatile uint8_t* p = td; for (int i = 0; i < 192; i++) p[i] = 0; }
        uint8_t* a = (uint8_t*)malloc(pk.size() ? pk.size() : 1);
        uint8_t* b = (uint8_t*)malloc(vk.size() ? vk.size() : 1);
        if (!a || !b) { free(a); free(b); throw std::bad_alloc(); }
        memcpy(a, pk.data(), pk.size());
        memcpy(b, vk.data(), vk.size());
        *pk_out = a; *pk_len = pk.size();
        *vk_out = b; *vk_len = vk.size();
    });
}

// ------------------------------------------------------------------------------------------------ verifier
int g16_verify_init(const uint8_t* vk, size_t vk_len, int device, g16_vctx** out) {
    return guarded([&] {
        REQUIRE(vk && vk_len && out, "NULL argument");
        *out = nullptr;
        require_device();
        std::unique_ptr<g16_vctx> c(new g16_vctx());
        c->v = vctx_create(vk, vk_len, device);
        *out = c.release();
    });
}
void g16_verify_free(g16_vctx* ctx) {
    if (!ctx) return;
    try {
        cudaSetDevice(ctx->v->device);
        delete ctx;
    } catch (...) {
    }
}
int g16_verify_info(const g16_vctx* ctx, uint64_t info[4]) {
    return guarded([&] {
        REQUIRE(ctx && info, "NULL argument");
        info[0] = ctx->v->n_public; info[1] = ctx->v->n_commit; info[2] = ctx->v->proof_bytes(); info[3] = ctx->v->nK;
    });
}
int g16_verify_batch(g16_vctx* ctx, size_t n, const uint8_t* proofs, const void* public_inputs, int public_format,
                     uint8_t* ok_out, float* device_ms) {
    return guarded([&] {
        REQUIRE(ctx && proofs && ok_out, "NULL argument");
        REQUIRE(public_inputs || ctx->v->n_public == 0, "NULL public inputs");
        REQUIRE(n > 0 && n <= (1u << 20), "batch size out of range");
        REQUIRE(public_format == 0 || public_format == 1, "public_format must be 0 (Montgomery limbs) or 1 (big-endian bytes)");
        std::lock_guard<std::mutex> lk(ctx->mu);
        float ms = vctx_verify_batch(*ctx->v, n, proofs, public_inputs, public_format, ok_out);
        if (device_ms) *device_ms = ms;
    });
}

// ------------------------------------------------------------------------------------------------ stage-level: pairing
int g16_pairing_check(const uint64_t* g1_points, const uint64_t* g2_points, size_t pairs_per_check, size_t n_checks,
                      uint8_t* ok_out) {
    return guarded([&] {
        REQUIRE(g1_points && g2_points && ok_out, "NULL argument");
        REQUIRE(pairs_per_check >= 1 && pairs_per_check <= 64 && n_checks >= 1 && pairs_per_check * n_checks <= (1u << 20),
                "pairing check: sizes out of range");
        require_device();
        const size_t np = pairs_per_check * n_checks;
        DevBuf<G1Affine> P;
        DevBuf<G2Affine> Q;
        DevBuf<uint8_t> ok(n_checks);
        P.upload((const G1Affine*)g1_points, np);
        Q.upload((const G2Affine*)g2_points, np);
        PairingWorkspace ws;
        pairing_check_run(ws, P.p, Q.p, (uint32_t)pairs_per_check, (uint32_t)n_checks, ok.p, 0);
        ok.download(ok_out, n_checks);
        G16_CUDA(cudaStreamSynchronize(0));
    });
}

int g16_g2_subgroup_check(const uint64_t* g2_points, size_t n, uint8_t* ok_out) {
    return guarded([&] {
        REQUIRE(g2_points && ok_out && n >= 1 && n <= (1u << 20), "bad argument");
        require_device();
        DevBuf<G2Affine> Q;
        DevBuf<uint8_t> ok(n);
        Q.upload((const G2Affine*)g2_points, n);
        PairingWorkspace ws;
        pairing_consts_ensure(ws, 0);
        launch_g2_subgroup(Q.p, (uint32_t)n, ws.consts.p, ok.p, 0);
        ok.download(ok_out, n);
        G16_CUDA(cudaStreamSynchronize(0));
    });
}

// ------------------------------------------------------------------------------------------------ stage-level: NTT
int g16_ntt(uint64_t* data, size_t n, int inverse, int coset, float* ms) {
    return guarded([&] {
        REQUIRE(data && n >= 2 && n <= (1u << 26), "bad argument");
        require_device();
        int k = log2_exact(n);
        NttDomain d;
        ntt_standalone_domain(d, k, 0);
        DevBuf<Fr> a, t(n);
        a.upload((const Fr*)data, n);
        cudaEvent_t e0, e1;
        G16_CUDA(cudaEventCreate(&e0));
        G16_CUDA(cudaEventCreate(&e1));
        G16_CUDA(cudaEventRecord(e0, 0));
        if (!inverse) {
            // natural -> (bit reversal) -> DIT with optional coset pre-scaling on the bit-reversed side
            ntt_bitrev(a.p, t.p, k, 0);
            ntt_run(d, t.p, n, 1, false, false, coset ? d.scale_coset_only.p : nullptr, 0);
        } else {
            // DIF with the inverse root, scaled on the bit-reversed side, then back to natural order
            ntt_run(d, a.p, n, 1, true, true, coset ? d.scale_coset_inv.p : d.scale_ninv.p, 0);
            ntt_bitrev(a.p, t.p, k, 0);
        }
        G16_CUDA(cudaEventRecord(e1, 0));
        t.download((Fr*)data, n);
        G16_CUDA(cudaDeviceSynchronize());
        if (ms) cudaEventElapsedTime(ms, e0, e1);
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    });
}

// forward DIF then inverse DIT (bit-reversed in between): a round trip that must reproduce the input exactly.
// checksum = number of mismatching elements after the last round trip (0 = exact).
int g16_ntt_bench(size_t n, size_t batch, int iters, float* ms_per_iter, uint64_t* checksum) {
    return guarded([&] {
        REQUIRE(n >= 2 && n <= (1u << 26) && batch >= 1 && iters >= 1 && ms_per_iter, "bad argument");
        require_device();
        int k = log2_exact(n);
        NttDomain d;
        ntt_standalone_domain(d, k, 0);
        size_t total = n * batch;
        DevBuf<Fr> a(total), ref(total);
        DevBuf<uint32_t> mism(1);
        mism.zero();
        ntt_fill_pattern(ref.p, total, 0);
        G16_CUDA(cudaMemcpy(a.p, ref.p, total * sizeof(Fr), cudaMemcpyDeviceToDevice));
        cudaEvent_t e0, e1;
        G16_CUDA(cudaEventCreate(&e0));
        G16_CUDA(cudaEventCreate(&e1));
        // warm-up round trip
        ntt_run(d, a.p, n, (uint32_t)batch, true, false, nullptr, 0);
        ntt_run(d, a.p, n, (uint32_t)batch, false, true, d.scale_ninv.p, 0);
        G16_CUDA(cudaEventRecord(e0, 0));
        for (int it = 0; it < iters; it++) {
            ntt_run(d, a.p, n, (uint32_t)batch, true, false, nullptr, 0);
            ntt_run(d, a.p, n, (uint32_t)batch, false, true, d.scale_ninv.p, 0);
        }
        G16_CUDA(cudaEventRecord(e1, 0));
        ntt_count_mismatches(a.p, ref.p, total, mism.p, 0);
        uint32_t m = 0;
        mism.download(&m, 1);
        G16_CUDA(cudaDeviceSynchronize());
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        *ms_per_iter = ms / (float)(2 * iters);   // per transform
        if (checksum) *checksum = m;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    });
}

int g16_imad_chain_rate(double* wide_carry_mads_per_s) {
    return guarded([&] {
        REQUIRE(wide_carry_mads_per_s, "NULL argument");
        *wide_carry_mads_per_s = imad_chain_rate();
    });
}
int g16_imad_peak(double* imad_per_s, double* imad_wide_per_s, double* modmul_per_s) {
    return guarded([&] {
        REQUIRE(imad_per_s && imad_wide_per_s && modmul_per_s, "NULL argument");
        require_device();
        imad_peak_measure(imad_per_s, imad_wide_per_s, modmul_per_s);
    });
}

}  // extern "C"
