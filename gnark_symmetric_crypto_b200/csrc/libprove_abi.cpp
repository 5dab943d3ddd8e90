// Outer ABI, verifier side: `Verify` of libraries/verifier/libverify.go:14-17 (+ InitVerifier, see below).
// Outer ABI: the four cgo exports of the reference's c-shared library libraries/prover/libprove.go
// (enforce_binding :17-18, InitAlgorithm :20-23, Free :25-28, Prove :30-47) and the JSON layer of
// libraries/prover/impl/prove_impl.go:116-143 / provers.go:53-59,79-89 restated in C++ (no Go toolchain on this box),
// on top of the g16_* inner seam. Same names, argument meaning and error behaviour:
//   * InitAlgorithm returns 0 and prints the reason for unknown ids / unparsable keys (prove_impl.go:65-114)
//   * Prove has no status channel: failures ("panics" in the reference) come back AS the payload, JSON-encoded
//     (libprove.go:33-43); success is {"proof":{"proofJson":<base64>},"publicSignals":<base64>} (prove_impl.go:129-134)
//   * the result buffer is malloc'd and must be released with Free (libprove.go:25-28,40,46)
#include "../../include/g16b200.h"
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

const char* const ALG_NAMES[3] = {"chacha20", "aes-128-ctr", "aes-256-ctr"};   // prove_impl.go:15-25
g16_ctx* g_provers[3] = {nullptr, nullptr, nullptr};
g16_vctx* g_verifiers[3] = {nullptr, nullptr, nullptr};   // libraries/verifier/impl/verify_impl.go:24 `verifiers`
std::mutex g_mu;

struct Panic : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---- minimal JSON reader for InputParams (provers.go:53-59). []uint8 fields accept an array of numbers or a base64
//      string, as Go's encoding/json does.
struct Json {
    const char* p;
    const char* end;
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
    bool eat(char c) { ws(); if (p < end && *p == c) { p++; return true; } return false; }
    void expect(char c) { if (!eat(c)) throw Panic(std::string("invalid character in JSON input, expected '") + c + "'"); }
    std::string str() {
        ws();
        if (p >= end || *p != '"') throw Panic("json: expected string");
        p++;
        std::string s;
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) {
                p++;
                switch (*p) {
                    case 'n': s += '\n'; break; case 't': s += '\t'; break; case 'r': s += '\r'; break;
                    case 'b': s += '\b'; break; case 'f': s += '\f'; break;
                    case 'u': { if (p + 4 < end) { unsigned v = (unsigned)strtoul(std::string(p + 1, 4).c_str(), nullptr, 16); s += (char)v; p += 4; } break; }
                    default: s += *p;
                }
                p++;
            } else {
                s += *p++;
            }
        }
        if (p >= end) throw Panic("unexpected end of JSON input");
        p++;
        return s;
    }
    double num() {
        ws();
        char* e = nullptr;
        double v = strtod(p, &e);
        if (e == p) throw Panic("json: expected number");
        p = e;
        return v;
    }
    void skip_value() {
        ws();
        if (p >= end) throw Panic("unexpected end of JSON input");
        if (*p == '"') { str(); return; }
        if (*p == '{') { p++; if (eat('}')) return; do { str(); expect(':'); skip_value(); } while (eat(',')); expect('}'); return; }
        if (*p == '[') { p++; if (eat(']')) return; do { skip_value(); } while (eat(',')); expect(']'); return; }
        if (!strncmp(p, "true", 4)) { p += 4; return; }
        if (!strncmp(p, "false", 5)) { p += 5; return; }
        if (!strncmp(p, "null", 4)) { p += 4; return; }
        num();
    }
};

int b64val(char c) {
    if (c >= 'A' && c <= 'Z') return c - 'A';
    if (c >= 'a' && c <= 'z') return c - 'a' + 26;
    if (c >= '0' && c <= '9') return c - '0' + 52;
    if (c == '+') return 62;
    if (c == '/') return 63;
    return -1;
}
std::vector<uint8_t> b64decode(const std::string& s) {
    std::vector<uint8_t> out;
    uint32_t acc = 0;
    int bits = 0;
    for (char c : s) {
        if (c == '=') break;
        int v = b64val(c);
        if (v < 0) throw Panic("illegal base64 data");
        acc = (acc << 6) | (uint32_t)v;
        bits += 6;
        if (bits >= 8) { bits -= 8; out.push_back((uint8_t)(acc >> bits)); }
    }
    return out;
}
std::string b64encode(const uint8_t* d, size_t n) {
    static const char T[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    std::string s;
    for (size_t i = 0; i < n; i += 3) {
        uint32_t v = (uint32_t)d[i] << 16;
        if (i + 1 < n) v |= (uint32_t)d[i + 1] << 8;
        if (i + 2 < n) v |= d[i + 2];
        s += T[(v >> 18) & 63];
        s += T[(v >> 12) & 63];
        s += i + 1 < n ? T[(v >> 6) & 63] : '=';
        s += i + 2 < n ? T[v & 63] : '=';
    }
    return s;
}
std::vector<uint8_t> read_bytes(Json& j) {
    j.ws();
    if (j.p < j.end && *j.p == '"') return b64decode(j.str());
    if (j.p + 4 <= j.end && !strncmp(j.p, "null", 4)) { j.p += 4; return {}; }
    std::vector<uint8_t> out;
    j.expect('[');
    if (j.eat(']')) return out;
    do {
        double v = j.num();
        if (v < 0 || v > 255 || v != (double)(int)v) throw Panic("json: cannot unmarshal number into Go value of type uint8");
        out.push_back((uint8_t)v);
    } while (j.eat(','));
    j.expect(']');
    return out;
}

struct InputParams {
    std::string cipher;
    std::vector<uint8_t> key, nonce, input;
    uint32_t counter = 0;
};
InputParams parse_params(const uint8_t* data, size_t len) {
    Json j{(const char*)data, (const char*)data + len};
    InputParams ip;
    j.expect('{');
    if (!j.eat('}')) {
        do {
            std::string k = j.str();
            j.expect(':');
            if (k == "cipher") ip.cipher = j.str();
            else if (k == "key") ip.key = read_bytes(j);
            else if (k == "nonce") ip.nonce = read_bytes(j);
            else if (k == "input") ip.input = read_bytes(j);
            else if (k == "counter") {
                j.ws();
                if (j.p < j.end && (*j.p == '[' || *j.p == '"' || *j.p == '{'))
                    throw Panic("json: cannot unmarshal into Go struct field InputParams.counter of type uint32");
                double v = j.num();
                if (v < 0 || v > 4294967295.0 || v != (double)(uint64_t)v) throw Panic("json: cannot unmarshal number into Go struct field InputParams.counter of type uint32");
                ip.counter = (uint32_t)v;
            } else j.skip_value();
        } while (j.eat(','));
        j.expect('}');
    }
    return ip;
}

std::string json_string(const std::string& s) {
    std::string o = "\"";
    for (char c : s) {
        if (c == '"' || c == '\\') { o += '\\'; o += c; }
        else if (c == '\n') o += "\\n";
        else if ((unsigned char)c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; }
        else o += c;
    }
    return o + "\"";
}

// ---- dynamic batching. The reference's API proves one request per call and its callers (the attestor) issue calls
// concurrently; the GPU only reaches its throughput on batches. Concurrent Prove calls for one cipher are therefore coalesced:
// each call enqueues its request and sleeps, one worker thread per cipher takes everything that is queued (up to
// G16_BATCH_MAX, default 1024) whenever the previous batch has finished, proves it in one g16_prove_*_batch call and wakes
// the callers. A lone request is picked up immediately, so its latency is unchanged. G16_DYNAMIC_BATCH=0 disables it.
struct Pending {
    const InputParams* ip = nullptr;
    uint8_t proof[196];
    uint8_t ct[64];
    int rc = 0;
    std::string err;
    bool done = false;
};
struct Batcher {
    int alg = 0;
    g16_ctx* ctx = nullptr;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<Pending*> q;
    bool stop = false;
    std::thread worker;
    size_t max_batch = 1024;
    size_t proof_bytes() const { return alg == 0 ? 164 : 196; }

    int prove_many(std::vector<Pending*>& b) {
        const size_t n = b.size();
        const size_t klen = b[0]->ip->key.size();
        for (Pending* p : b)
            if (p->ip->key.size() != klen) return G16_ERR_ARG;   // mixed key lengths: the caller falls back to single requests
        std::vector<uint8_t> keys(klen * n), nonces(12 * n), inputs(64 * n), proofs(proof_bytes() * n), cts(64 * n);
        std::vector<uint32_t> counters(n);
        for (size_t i = 0; i < n; i++) {
            memcpy(&keys[klen * i], b[i]->ip->key.data(), klen);
            memcpy(&nonces[12 * i], b[i]->ip->nonce.data(), 12);
            memcpy(&inputs[64 * i], b[i]->ip->input.data(), 64);
            counters[i] = b[i]->ip->counter;
        }
        int rc = alg == 0 ? g16_prove_chacha_batch(ctx, n, keys.data(), nonces.data(), counters.data(), inputs.data(), nullptr,
                                                   proofs.data(), cts.data())
                          : g16_prove_aes_batch(ctx, n, keys.data(), klen, nonces.data(), counters.data(), inputs.data(), nullptr,
                                                proofs.data(), cts.data());
        if (rc) return rc;
        for (size_t i = 0; i < n; i++) {
            memcpy(b[i]->proof, &proofs[proof_bytes() * i], proof_bytes());
            memcpy(b[i]->ct, &cts[64 * i], 64);
            b[i]->rc = 0;
        }
        return 0;
    }
    void run() {
        for (;;) {
            std::vector<Pending*> batch;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stop || !q.empty(); });
                if (stop && q.empty()) return;
                while (!q.empty() && batch.size() < max_batch) { batch.push_back(q.front()); q.pop_front(); }
            }
            int rc = prove_many(batch);
            if (rc) {
                const std::string msg = g16_last_error();
                if (batch.size() == 1) {
                    batch[0]->rc = rc;
                    batch[0]->err = msg;
                } else {
                    // one request of the batch cannot be proved (e.g. an AES counter the circuit rejects): prove them one by
                    // one so that only the offending call fails
                    for (Pending* p : batch) {
                        std::vector<Pending*> one(1, p);
                        p->rc = prove_many(one);
                        if (p->rc) p->err = g16_last_error();
                    }
                }
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                for (Pending* p : batch) p->done = true;
            }
            cv_done.notify_all();
        }
    }
};
Batcher* g_batchers[3] = {nullptr, nullptr, nullptr};

std::string prove_impl(const uint8_t* params, size_t len) {
    InputParams ip = parse_params(params, len);
    int alg = -1;
    for (int i = 0; i < 3; i++) if (ip.cipher == ALG_NAMES[i]) alg = i;
    if (alg < 0) throw Panic("could not find prover for" + ip.cipher);   // prove_impl.go:140-142 (sic, no space)
    g16_ctx* ctx;
    Batcher* bt;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        ctx = g_provers[alg];
        bt = g_batchers[alg];
    }
    if (!ctx) throw Panic("proving params are not initialized for cipher: " + ip.cipher);   // prove_impl.go:124-126
    if (alg == 0) {
        if (ip.key.size() != 32) throw Panic("key length must be 32: " + std::to_string(ip.key.size()));         // provers.go:81-83
        if (ip.nonce.size() != 12) throw Panic("nonce length must be 12: " + std::to_string(ip.nonce.size()));   // :84-86
        if (ip.input.size() != 64) throw Panic("plaintext length must be 64: " + std::to_string(ip.input.size()));   // :87-89
    } else {
        if (ip.key.size() != 32 && ip.key.size() != 16) throw Panic("key length must be 16 or 32: " + std::to_string(ip.key.size()));   // provers.go:174-176
        if (ip.nonce.size() != 12) throw Panic("nonce length must be 12: " + std::to_string(ip.nonce.size()));
        if (ip.input.size() != 64) throw Panic("plaintext length must be 64: " + std::to_string(ip.input.size()));
    }
    const size_t pb = alg == 0 ? 164 : 196;
    Pending pd;
    pd.ip = &ip;
    if (bt) {
        std::unique_lock<std::mutex> lk(bt->mu);
        bt->q.push_back(&pd);
        bt->cv_work.notify_one();
        bt->cv_done.wait(lk, [&] { return pd.done; });
    } else {
        int rc = alg == 0 ? g16_prove_chacha_batch(ctx, 1, ip.key.data(), ip.nonce.data(), &ip.counter, ip.input.data(), nullptr, pd.proof, pd.ct)
                          : g16_prove_aes_batch(ctx, 1, ip.key.data(), ip.key.size(), ip.nonce.data(), &ip.counter, ip.input.data(),
                                                nullptr, pd.proof, pd.ct);
        pd.rc = rc;
        if (rc) pd.err = g16_last_error();
    }
    if (pd.rc) throw Panic("groth16 prove failed: " + pd.err);
    return "{\"proof\":{\"proofJson\":\"" + b64encode(pd.proof, pb) + "\"},\"publicSignals\":\"" + b64encode(pd.ct, 64) + "\"}";
}

// ---- libraries/verifier: InputVerifyParams (verify_impl.go:18-22) and the public-witness layouts of verifiers.go:50-152
struct VerifyParams {
    std::string cipher;
    std::vector<uint8_t> proof, signals;
};
VerifyParams parse_verify_params(const uint8_t* data, size_t len) {
    Json j{(const char*)data, (const char*)data + len};
    VerifyParams vp;
    j.expect('{');
    if (!j.eat('}')) {
        do {
            std::string k = j.str();
            j.expect(':');
            if (k == "cipher") vp.cipher = j.str();
            else if (k == "proof") vp.proof = read_bytes(j);
            else if (k == "publicSignals") vp.signals = read_bytes(j);
            else j.skip_value();
        } while (j.eat(','));
        j.expect('}');
    }
    return vp;
}
void push_be32(std::vector<uint8_t>& out, uint32_t v) {   // one field element, 32-byte big-endian
    out.insert(out.end(), 28, 0);
    out.push_back((uint8_t)(v >> 24)); out.push_back((uint8_t)(v >> 16)); out.push_back((uint8_t)(v >> 8)); out.push_back((uint8_t)v);
}
void push_word_bits(std::vector<uint8_t>& out, uint32_t w) {   // utils.Uint32ToBits: LSB first
    for (int b = 0; b < 32; b++) push_be32(out, (w >> b) & 1u);
}
uint32_t rd_le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint32_t rd_be32(const uint8_t* p) { return (uint32_t)p[3] | ((uint32_t)p[2] << 8) | ((uint32_t)p[1] << 16) | ((uint32_t)p[0] << 24); }

// concurrent Verify calls are coalesced the same way as Prove calls (one g16_verify_batch per drained queue)
struct VPending {
    const uint8_t* proof = nullptr;
    const uint8_t* pub = nullptr;
    uint8_t ok = 0;
    int rc = 0;
    bool done = false;
};
struct VBatcher {
    g16_vctx* ctx = nullptr;
    size_t proof_bytes = 0, pub_bytes = 0;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<VPending*> q;
    bool stop = false;
    std::thread worker;
    size_t max_batch = 4096;
    void run() {
        for (;;) {
            std::vector<VPending*> batch;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stop || !q.empty(); });
                if (stop && q.empty()) return;
                while (!q.empty() && batch.size() < max_batch) { batch.push_back(q.front()); q.pop_front(); }
            }
            const size_t n = batch.size();
            std::vector<uint8_t> proofs(proof_bytes * n), pubs(pub_bytes * n), ok(n, 0);
            for (size_t i = 0; i < n; i++) {
                memcpy(&proofs[proof_bytes * i], batch[i]->proof, proof_bytes);
                memcpy(&pubs[pub_bytes * i], batch[i]->pub, pub_bytes);
            }
            int rc = g16_verify_batch(ctx, n, proofs.data(), pubs.data(), 1, ok.data(), nullptr);
            {
                std::lock_guard<std::mutex> lk(mu);
                for (size_t i = 0; i < n; i++) { batch[i]->rc = rc; batch[i]->ok = rc ? 0 : ok[i]; batch[i]->done = true; }
            }
            cv_done.notify_all();
        }
    }
};
VBatcher* g_vbatchers[3] = {nullptr, nullptr, nullptr};

bool verify_impl(const uint8_t* params, size_t len) {
    VerifyParams vp = parse_verify_params(params, len);
    int alg = -1;
    for (int i = 0; i < 3; i++) if (vp.cipher == ALG_NAMES[i]) alg = i;
    if (alg < 0) return false;   // verify_impl.go:78-81
    g16_vctx* ctx;
    VBatcher* vb;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        ctx = g_verifiers[alg];
        vb = g_vbatchers[alg];
    }
    if (!ctx) { printf("verifying key is not initialized for cipher: %s\n", vp.cipher.c_str()); return false; }
    if (vp.signals.size() != 128 + 12 + 4) {   // verifiers.go:52-55,105-107: ciphertext | nonce | counter | plaintext
        printf("public signals must be 144 bytes, not %zu\n", vp.signals.size());
        return false;
    }
    uint64_t info[4];
    if (g16_verify_info(ctx, info)) return false;
    if (vp.proof.size() < info[2]) { printf("unexpected EOF\n"); return false; }   // Proof.ReadFrom on a short buffer
    const uint8_t* ct = vp.signals.data();
    const uint8_t* nonce = ct + 64;
    const uint8_t* counter = nonce + 12;
    const uint8_t* pt = counter + 4;
    std::vector<uint8_t> pub;
    pub.reserve((size_t)info[0] * 32);
    if (alg == 0) {
        // ChaChaCircuit public fields in declaration order: Counter, Nonce[3], In[16], Out[16] (words as LSB-first bits);
        // In / Out words big-endian, Nonce / Counter little-endian (verifiers.go:59-85, utils/bytes.go:11-47)
        push_word_bits(pub, rd_le32(counter));
        for (int k = 0; k < 3; k++) push_word_bits(pub, rd_le32(nonce + 4 * k));
        for (int k = 0; k < 16; k++) push_word_bits(pub, rd_be32(pt + 4 * k));
        for (int k = 0; k < 16; k++) push_word_bits(pub, rd_be32(ct + 4 * k));
    } else {
        // AESWrapper public fields: Nonce[12], Counter (big-endian u32), Plaintext[64], Ciphertext[64] (verifiers.go:109-127)
        for (int k = 0; k < 12; k++) push_be32(pub, nonce[k]);
        push_be32(pub, rd_be32(counter));
        for (int k = 0; k < 64; k++) push_be32(pub, pt[k]);
        for (int k = 0; k < 64; k++) push_be32(pub, ct[k]);
    }
    if (pub.size() != (size_t)info[0] * 32) { printf("verifying key does not match cipher %s\n", vp.cipher.c_str()); return false; }
    if (vb) {
        VPending pd;
        pd.proof = vp.proof.data();
        pd.pub = pub.data();
        std::unique_lock<std::mutex> lk(vb->mu);
        vb->q.push_back(&pd);
        vb->cv_work.notify_one();
        vb->cv_done.wait(lk, [&] { return pd.done; });
        return pd.rc == 0 && pd.ok != 0;
    }
    uint8_t ok = 0;
    int rc = g16_verify_batch(ctx, 1, vp.proof.data(), pub.data(), 1, &ok, nullptr);
    if (rc) { printf("%s\n", g16_last_error()); return false; }
    return ok != 0;
}

}  // namespace

extern "C" {

void enforce_binding(void) {}

// The reference embeds its three verifying keys with go:embed (verify_impl.go:26-33) and parses them in init() (:35-62); a C
// library receives the same bytes once per cipher instead.
unsigned char InitVerifier(unsigned char algorithmID, GoSlice_g16 verifyingKey) {
    if (algorithmID > 2) return 0;
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_verifiers[algorithmID]) return 1;
    if (!verifyingKey.data || verifyingKey.len <= 0) {
        printf("error reading verifying key: empty input\n");
        return 0;
    }
    int device = 0;
    if (const char* d = getenv("G16_DEVICE")) device = atoi(d);
    g16_vctx* ctx = nullptr;
    int rc = g16_verify_init((const uint8_t*)verifyingKey.data, (size_t)verifyingKey.len, device, &ctx);
    if (rc) {
        printf("error reading verifying key: %s\n", g16_last_error());
        return 0;
    }
    g_verifiers[algorithmID] = ctx;
    const char* dyn = getenv("G16_DYNAMIC_BATCH");
    uint64_t info[4];
    if ((!dyn || atoi(dyn) != 0) && g16_verify_info(ctx, info) == 0) {
        VBatcher* b = new VBatcher();
        b->ctx = ctx;
        b->proof_bytes = (size_t)info[2];
        b->pub_bytes = (size_t)info[0] * 32;
        b->worker = std::thread([b] { b->run(); });
        g_vbatchers[algorithmID] = b;
    }
    return 1;
}

// libraries/verifier/libverify.go:14-17 + impl/verify_impl.go:64-82: any failure (bad JSON, unknown cipher, malformed proof,
// a recovered panic) is `false`
unsigned char Verify(GoSlice_g16 params) {
    try {
        if (!params.data || params.len <= 0) return 0;
        return verify_impl((const uint8_t*)params.data, (size_t)params.len) ? 1 : 0;
    } catch (const std::exception& e) {
        printf("%s\n", e.what());
        return 0;
    }
}

unsigned char InitAlgorithm(unsigned char algorithmID, GoSlice_g16 provingKey, GoSlice_g16 r1cs) {
    if (algorithmID > 2) return 0;   // prove_impl.go:72,113
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_provers[algorithmID]) return 1;   // initDone, prove_impl.go:74-76
    if (!provingKey.data || !r1cs.data || provingKey.len < 0 || r1cs.len < 0) {
        printf("error reading proving key: empty input\n");
        return 0;
    }
    int device = 0;
    if (const char* d = getenv("G16_DEVICE")) device = atoi(d);
    g16_ctx* ctx = nullptr;
    int rc = g16_init((const uint8_t*)provingKey.data, (size_t)provingKey.len, (const uint8_t*)r1cs.data, (size_t)r1cs.len,
                      device, &ctx);
    if (rc) {
        printf("error reading proving key: %s\n", g16_last_error());   // prove_impl.go:88-91,104-107
        return 0;
    }
    g_provers[algorithmID] = ctx;
    const char* dyn = getenv("G16_DYNAMIC_BATCH");
    if (!dyn || atoi(dyn) != 0) {
        Batcher* b = new Batcher();
        b->alg = algorithmID;
        b->ctx = ctx;
        if (const char* m = getenv("G16_BATCH_MAX")) { int v = atoi(m); if (v > 0 && v <= 65536) b->max_batch = (size_t)v; }
#if !defined(G16_EMU)
        if (algorithmID == 0) {
            // Warm start. Every (key, nonce, counter, input) is a valid ChaCha request, so one dummy batch of the largest size
            // the worker can form builds the key tables (both bases of the Z query), sizes every device buffer and touches
            // every stage before the first caller arrives; without it a serving process keeps meeting first-time stalls
            // (table build, buffer growth) while the batch sizes it sees still grow. G16_PREWARM=0 skips it, =N sets the size.
            size_t n = b->max_batch;
            if (const char* w = getenv("G16_PREWARM")) n = (size_t)(atoi(w) > 0 ? atoi(w) : 0);
            if (n) {
                std::vector<uint8_t> keys(n * 32), nonces(n * 12), inputs(n * 64), proofs(n * 164), cts(n * 64);
                std::vector<uint32_t> counters(n);
                if (g16_prove_chacha_batch(ctx, n, keys.data(), nonces.data(), counters.data(), inputs.data(), nullptr,
                                           proofs.data(), cts.data()))
                    printf("warm-up batch failed: %s\n", g16_last_error());
            }
        }
#endif
        b->worker = std::thread([b] { b->run(); });
        g_batchers[algorithmID] = b;
    }
    return 1;
}

void Free(void* pointer) { free(pointer); }

Prove_return_g16 Prove(GoSlice_g16 params) {
    std::string res;
    try {
        if (!params.data || params.len <= 0) throw Panic("unexpected end of JSON input");
        res = prove_impl((const uint8_t*)params.data, (size_t)params.len);
    } catch (const std::exception& e) {
        printf("%s\n", e.what());          // libprove.go:35
        res = json_string(e.what());       // libprove.go:36-41: json.Marshal(err) returned as the payload
    }
    Prove_return_g16 r;
    r.r1 = (long long)res.size();
    r.r0 = malloc(res.size() ? res.size() : 1);   // C.CBytes, libprove.go:40,46
    if (r.r0) memcpy(r.r0, res.data(), res.size());
    else r.r1 = 0;
    return r;
}

// test hook: drop the cached provers (the reference has no such call; its map lives for the process lifetime)
void g16_libprove_reset(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& b : g_batchers) {
        if (!b) continue;
        { std::lock_guard<std::mutex> lk2(b->mu); b->stop = true; }
        b->cv_work.notify_all();
        if (b->worker.joinable()) b->worker.join();
        delete b;
        b = nullptr;
    }
    for (auto& p : g_provers) { if (p) g16_free(p); p = nullptr; }
    for (auto& b : g_vbatchers) {
        if (!b) continue;
        { std::lock_guard<std::mutex> lk2(b->mu); b->stop = true; }
        b->cv_work.notify_all();
        if (b->worker.joinable()) b->worker.join();
        delete b;
        b = nullptr;
    }
    for (auto& v : g_verifiers) { if (v) g16_verify_free(v); v = nullptr; }
}

}  // extern "C"
