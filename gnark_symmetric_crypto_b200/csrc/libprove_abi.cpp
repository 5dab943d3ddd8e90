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

// A "panic" of the reference. libprove.go:33-43 returns json.Marshal(recovered value) as the payload: a string panic
// (log.Panicf, panic(fmt.Sprintf(..))) marshals as a JSON string; a panic(err) marshals the error VALUE, i.e. an object with
// the error type's exported fields ({"Offset":N} for *json.SyntaxError, {} for most others) — `payload` holds that form.
struct Panic : std::runtime_error {
    std::string payload;   // empty: marshal what() as a JSON string
    explicit Panic(const std::string& msg, std::string obj = std::string()) : std::runtime_error(msg), payload(std::move(obj)) {}
};
[[noreturn]] void syntax_error(const std::string& msg, size_t offset) {   // *json.SyntaxError{msg (unexported), Offset}
    throw Panic(msg, "{\"Offset\":" + std::to_string(offset) + "}");
}
std::string json_string(const std::string& s);
[[noreturn]] void type_error(const std::string& value, const std::string& field, const std::string& go_type, size_t offset) {
    // *json.UnmarshalTypeError{Value, Type reflect.Type, Offset, Struct, Field}; a reflect.Type marshals as {}
    throw Panic("json: cannot unmarshal " + value + " into Go struct field InputParams." + field + " of type " + go_type,
                "{\"Value\":" + json_string(value) + ",\"Type\":{},\"Offset\":" + std::to_string(offset) +
                    ",\"Struct\":\"InputParams\",\"Field\":" + json_string(field) + "}");
}

// ---- JSON reader for InputParams (provers.go:53-59), bounded by [p, end) (the GoSlice is not NUL-terminated). []uint8 fields
//      accept a base64 string, null, or an array of numbers, as Go's encoding/json does; unsigned fields accept only the
//      integer grammar (encoding/json hands the literal to strconv.ParseUint: "1e3", "1.0", "-1", "0x10" are type errors).
struct Json {
    const char* base;
    const char* p;
    const char* end;
    size_t off() const { return (size_t)(p - base); }
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
    bool eat(char c) { ws(); if (p < end && *p == c) { p++; return true; } return false; }
    void expect(char c) {
        ws();
        if (p >= end) syntax_error("unexpected end of JSON input", off());
        if (*p != c) syntax_error(std::string("invalid character '") + *p + "' looking for '" + c + "'", off() + 1);
        p++;
    }
    bool lit(const char* w) {
        size_t n = strlen(w);
        if ((size_t)(end - p) >= n && !memcmp(p, w, n)) { p += n; return true; }
        return false;
    }
    static void utf8(std::string& s, uint32_t cp) {
        if (cp < 0x80) s += (char)cp;
        else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 63)); }
        else if (cp < 0x10000) { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 63)); s += (char)(0x80 | (cp & 63)); }
        else { s += (char)(0xF0 | (cp >> 18)); s += (char)(0x80 | ((cp >> 12) & 63)); s += (char)(0x80 | ((cp >> 6) & 63)); s += (char)(0x80 | (cp & 63)); }
    }
    bool hex4(uint32_t& v) {
        if (end - p < 4) return false;
        v = 0;
        for (int k = 0; k < 4; k++) {
            char c = p[k];
            int d = c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1;
            if (d < 0) return false;
            v = v * 16 + (uint32_t)d;
        }
        p += 4;
        return true;
    }
    std::string str() {
        ws();
        if (p >= end) syntax_error("unexpected end of JSON input", off());
        if (*p != '"') syntax_error(std::string("invalid character '") + *p + "' looking for beginning of string", off() + 1);
        p++;
        std::string s;
        while (p < end && *p != '"') {
            if ((unsigned char)*p < 0x20) syntax_error("invalid character in string literal", off() + 1);
            if (*p == '\\') {
                p++;
                if (p >= end) break;
                char c = *p++;
                switch (c) {
                    case 'n': s += '\n'; break; case 't': s += '\t'; break; case 'r': s += '\r'; break;
                    case 'b': s += '\b'; break; case 'f': s += '\f'; break;
                    case '"': case '\\': case '/': s += c; break;
                    case 'u': {
                        uint32_t v;
                        if (!hex4(v)) syntax_error("invalid character in \\u hexadecimal character escape", off() + 1);
                        if (v >= 0xD800 && v < 0xDC00 && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {   // surrogate pair
                            const char* save = p;
                            p += 2;
                            uint32_t lo;
                            if (hex4(lo) && lo >= 0xDC00 && lo < 0xE000) v = 0x10000 + ((v - 0xD800) << 10) + (lo - 0xDC00);
                            else { p = save; v = 0xFFFD; }
                        } else if (v >= 0xD800 && v < 0xE000) v = 0xFFFD;   // lone surrogate, as encoding/json
                        utf8(s, v);
                        break;
                    }
                    default: syntax_error(std::string("invalid character '") + c + "' in string escape code", off());
                }
            } else {
                s += *p++;
            }
        }
        if (p >= end) syntax_error("unexpected end of JSON input", off());
        p++;
        return s;
    }
    // the JSON number grammar, nothing else; returns the literal
    std::string number() {
        ws();
        const char* b = p;
        if (p < end && *p == '-') p++;
        if (p >= end) syntax_error("unexpected end of JSON input", off());
        if (*p == '0') p++;
        else if (*p >= '1' && *p <= '9') { while (p < end && *p >= '0' && *p <= '9') p++; }
        else syntax_error(std::string("invalid character '") + *p + "' looking for beginning of value", off() + 1);
        if (p < end && *p == '.') {
            p++;
            if (p >= end || *p < '0' || *p > '9') syntax_error("invalid character after decimal point in numeric literal", off() + 1);
            while (p < end && *p >= '0' && *p <= '9') p++;
        }
        if (p < end && (*p == 'e' || *p == 'E')) {
            p++;
            if (p < end && (*p == '+' || *p == '-')) p++;
            if (p >= end || *p < '0' || *p > '9') syntax_error("invalid character in exponent of numeric literal", off() + 1);
            while (p < end && *p >= '0' && *p <= '9') p++;
        }
        return std::string(b, p);
    }
    // unsigned integer field of `bits` bits (strconv.ParseUint on the literal)
    uint64_t uint_field(const std::string& field, const std::string& go_type, int bits) {
        std::string n = number();
        bool ok = !n.empty() && n.size() <= 20;
        uint64_t v = 0;
        for (char c : n) {
            if (c < '0' || c > '9') { ok = false; break; }
            v = v * 10 + (uint64_t)(c - '0');
        }
        if (!ok || (bits < 64 && (v >> bits))) type_error("number " + n, field, go_type, off());
        return v;
    }
    void skip_value() {
        ws();
        if (p >= end) syntax_error("unexpected end of JSON input", off());
        if (*p == '"') { str(); return; }
        if (*p == '{') { p++; if (eat('}')) return; do { str(); expect(':'); skip_value(); } while (eat(',')); expect('}'); return; }
        if (*p == '[') { p++; if (eat(']')) return; do { skip_value(); } while (eat(',')); expect(']'); return; }
        if (lit("true") || lit("false") || lit("null")) return;
        number();
    }
    const char* kind_name() {   // what encoding/json calls the value at p in a type error
        ws();
        if (p >= end) return "value";
        if (*p == '"') return "string";
        if (*p == '{') return "object";
        if (*p == '[') return "array";
        if (*p == 't' || *p == 'f') return "bool";
        return "number";
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
// base64.StdEncoding.DecodeString: padded alphabet, '\r' and '\n' ignored, anything else (missing or misplaced padding,
// trailing data) is a CorruptInputError (marshals as a bare number: its underlying type is int64)
std::vector<uint8_t> b64decode(const std::string& in) {
    std::string s;
    for (char c : in) if (c != '\r' && c != '\n') s += c;
    auto corrupt = [](size_t at) -> void { throw Panic("illegal base64 data at input byte " + std::to_string(at), std::to_string(at)); };
    if (s.size() % 4) corrupt(s.size() - s.size() % 4);
    std::vector<uint8_t> out;
    for (size_t i = 0; i < s.size(); i += 4) {
        int v[4];
        int pad = 0;
        for (int k = 0; k < 4; k++) {
            char c = s[i + k];
            if (c == '=') {
                if (i + 4 != s.size() || k < 2) corrupt(i + k);   // padding only in the last quantum, at most two
                pad++;
                v[k] = 0;
            } else {
                if (pad) corrupt(i + k);
                v[k] = b64val(c);
                if (v[k] < 0) corrupt(i + k);
            }
        }
        uint32_t acc = ((uint32_t)v[0] << 18) | ((uint32_t)v[1] << 12) | ((uint32_t)v[2] << 6) | (uint32_t)v[3];
        out.push_back((uint8_t)(acc >> 16));
        if (pad < 2) out.push_back((uint8_t)(acc >> 8));
        if (pad < 1) out.push_back((uint8_t)acc);
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
std::vector<uint8_t> read_bytes(Json& j, const std::string& field) {
    j.ws();
    if (j.p < j.end && *j.p == '"') return b64decode(j.str());
    if (j.lit("null")) return {};
    if (j.p < j.end && *j.p != '[') type_error(j.kind_name(), field, "[]uint8", j.off() + 1);
    std::vector<uint8_t> out;
    j.expect('[');
    if (j.eat(']')) return out;
    do {
        j.ws();
        if (j.p < j.end && (*j.p == '"' || *j.p == '{' || *j.p == '[' || *j.p == 't' || *j.p == 'f')) type_error(j.kind_name(), field, "uint8", j.off() + 1);
        if (j.lit("null")) { out.push_back(0); continue; }   // null leaves the element at its zero value
        out.push_back((uint8_t)j.uint_field(field, "uint8", 8));
    } while (j.eat(','));
    j.expect(']');
    return out;
}

bool key_is(const std::string& k, const char* name) {   // encoding/json matches field names case-insensitively
    size_t n = strlen(name);
    if (k.size() != n) return false;
    for (size_t i = 0; i < n; i++) if (tolower((unsigned char)k[i]) != tolower((unsigned char)name[i])) return false;
    return true;
}

struct InputParams {
    std::string cipher;
    std::vector<uint8_t> key, nonce, input;
    uint32_t counter = 0;
};
// one InputParams object at the cursor
InputParams parse_params_at(Json& j) {
    InputParams ip;
    j.ws();
    if (j.lit("null")) return ip;   // *InputParams stays nil in the reference; its field access then panics (caller checks cipher)
    if (j.p < j.end && *j.p != '{') {
        const char* kind = j.kind_name();
        throw Panic(std::string("json: cannot unmarshal ") + kind + " into Go value of type impl.InputParams",
                    std::string("{\"Value\":\"") + kind + "\",\"Type\":{},\"Offset\":" + std::to_string(j.off() + 1) + ",\"Struct\":\"\",\"Field\":\"\"}");
    }
    j.expect('{');
    if (!j.eat('}')) {
        do {
            std::string k = j.str();
            j.expect(':');
            if (key_is(k, "cipher")) {
                j.ws();
                if (j.lit("null")) continue;
                if (j.p < j.end && *j.p != '"') type_error(j.kind_name(), "cipher", "string", j.off() + 1);
                ip.cipher = j.str();
            } else if (key_is(k, "key")) ip.key = read_bytes(j, "key");
            else if (key_is(k, "nonce")) ip.nonce = read_bytes(j, "nonce");
            else if (key_is(k, "input")) ip.input = read_bytes(j, "input");
            else if (key_is(k, "counter")) {
                j.ws();
                if (j.lit("null")) continue;
                if (j.p < j.end && (*j.p == '[' || *j.p == '"' || *j.p == '{' || *j.p == 't' || *j.p == 'f'))
                    type_error(j.kind_name(), "counter", "uint32", j.off() + 1);
                ip.counter = (uint32_t)j.uint_field("counter", "uint32", 32);
            } else j.skip_value();
        } while (j.eat(','));
        j.expect('}');
    }
    return ip;
}
InputParams parse_params(const uint8_t* data, size_t len) {
    Json j{(const char*)data, (const char*)data, (const char*)data + len};
    InputParams ip = parse_params_at(j);
    j.ws();
    if (j.p < j.end) syntax_error(std::string("invalid character '") + *j.p + "' after top-level value", j.off() + 1);
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
std::string panic_payload(const std::exception& e) {
    if (const Panic* p = dynamic_cast<const Panic*>(&e))
        if (!p->payload.empty()) return p->payload;
    return json_string(e.what());
}

// ---- provers.go:79-89 / 172-182 input checks, plus what the circuit itself would reject, decided on the host so that a
// request that cannot be proved never enters a batch. Returns the algorithm id.
int validate_request(const InputParams& ip) {
    int alg = -1;
    for (int i = 0; i < 3; i++) if (ip.cipher == ALG_NAMES[i]) alg = i;
    if (alg < 0) throw Panic("could not find prover for" + ip.cipher);   // prove_impl.go:140-142 (sic, no space)
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!g_provers[alg]) throw Panic("proving params are not initialized for cipher: " + ip.cipher);   // prove_impl.go:124-126
    }
    if (alg == 0) {
        if (ip.key.size() != 32) throw Panic("key length must be 32: " + std::to_string(ip.key.size()));         // provers.go:81-83
        if (ip.nonce.size() != 12) throw Panic("nonce length must be 12: " + std::to_string(ip.nonce.size()));   // :84-86
        if (ip.input.size() != 64) throw Panic("plaintext length must be 64: " + std::to_string(ip.input.size()));   // :87-89
    } else {
        if (ip.key.size() != 32 && ip.key.size() != 16) throw Panic("key length must be 16 or 32: " + std::to_string(ip.key.size()));   // provers.go:174-176
        if (ip.nonce.size() != 12) throw Panic("nonce length must be 12: " + std::to_string(ip.nonce.size()));
        if (ip.input.size() != 64) throw Panic("plaintext length must be 64: " + std::to_string(ip.input.size()));
        // The reference gets past its own checks here and panics with an error VALUE further down (payload: a JSON object):
        // a key of the other AES size makes frontend.NewWitness / groth16.Prove fail against this cipher's circuit
        // (provers.go:212-219), and a counter above 0xFFFFFFFB violates AssertIsLessOrEqual(counter, MaxUint32) in one of the
        // four blocks (circuits/aesV2/aes128.go:50-53), so groth16.Prove returns a solver error.
        const size_t want = alg == 1 ? 16 : 32;
        if (ip.key.size() != want)
            throw Panic("witness does not match circuit " + ip.cipher + ": key length " + std::to_string(ip.key.size()), "{}");
        if (ip.counter > 0xFFFFFFFBu)
            throw Panic("groth16 prove failed: constraint is not satisfied: counter + 4 blocks exceeds 2^32 - 1", "{}");
    }
    return alg;
}

// ---- dynamic batching. The reference's API proves one request per call and its callers (the attestor) issue calls
// concurrently; the GPU only reaches its throughput on batches. Concurrent Prove calls for one cipher are therefore coalesced:
// each call enqueues its request and sleeps; one worker thread PER GPU (G16_DEVICES) takes its share of what is queued (up to
// G16_BATCH_MAX, default 1024) whenever its device is free, proves it in one g16_prove_*_batch call on that device and wakes
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
    std::vector<g16_ctx*> devs;   // borrowed single-device handles of the cipher's (multi-device) context
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::deque<Pending*> q;
    bool stop = false;
    std::vector<std::thread> workers;
    size_t idle = 0;              // workers waiting for work
    size_t max_batch = 1024;
    size_t batches = 0, proved = 0;   // statistics (g16_libprove_stats)
    size_t proof_bytes() const { return alg == 0 ? 164 : 196; }

    // One batch on one device. Per-request verdicts: a request the circuit rejects fails alone (g16_last_batch_status);
    // nothing is proved twice. Any other failure (CUDA error, out of memory) fails the whole batch once, with its message.
    void prove_many(g16_ctx* ctx, std::vector<Pending*>& b) {
        const size_t n = b.size();
        const size_t klen = b[0]->ip->key.size();   // uniform per cipher: validate_request
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
        const std::string msg = rc ? g16_last_error() : "";
        std::vector<uint32_t> st(n, 0);
        if (rc == G16_ERR_UNSAT && g16_last_batch_status(ctx, st.data(), n) != 0) st.assign(n, 1);
        for (size_t i = 0; i < n; i++) {
            const bool bad = rc == G16_ERR_UNSAT ? st[i] != 0 : rc != 0;
            b[i]->rc = bad ? rc : 0;
            if (bad) { b[i]->err = msg; continue; }
            memcpy(b[i]->proof, &proofs[proof_bytes() * i], proof_bytes());
            memcpy(b[i]->ct, &cts[64 * i], 64);
        }
        volatile uint8_t* kp = keys.data();   // cipher keys are secrets: no copies left on the heap
        for (size_t i = 0; i < keys.size(); i++) kp[i] = 0;
    }
    void run(size_t slot) {
        g16_ctx* ctx = devs[slot];
        for (;;) {
            std::vector<Pending*> batch;
            {
                std::unique_lock<std::mutex> lk(mu);
                idle++;
                cv_work.wait(lk, [&] { return stop || !q.empty(); });
                idle--;
                if (stop && q.empty()) return;
                // leave the other free devices their share of the queue: ceil(queued / (free workers + this one))
                size_t take = (q.size() + idle) / (idle + 1);
                if (take < 1) take = 1;
                if (take > max_batch) take = max_batch;
                while (!q.empty() && batch.size() < take) { batch.push_back(q.front()); q.pop_front(); }
                if (!q.empty() && idle) cv_work.notify_one();
            }
            prove_many(ctx, batch);
            {
                std::lock_guard<std::mutex> lk(mu);
                for (Pending* p : batch) p->done = true;
                batches++;
                proved += batch.size();
            }
            cv_done.notify_all();
        }
    }
};
Batcher* g_batchers[3] = {nullptr, nullptr, nullptr};

std::string output_json(const Pending& pd, size_t pb) {   // prove_impl.go:129-134
    return "{\"proof\":{\"proofJson\":\"" + b64encode(pd.proof, pb) + "\"},\"publicSignals\":\"" + b64encode(pd.ct, 64) + "\"}";
}

std::string prove_impl(const uint8_t* params, size_t len) {
    InputParams ip = parse_params(params, len);
    const int alg = validate_request(ip);
    g16_ctx* ctx;
    Batcher* bt;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        ctx = g_provers[alg];
        bt = g_batchers[alg];
    }
    if (!ctx) throw Panic("proving params are not initialized for cipher: " + ip.cipher);
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
    if (pd.rc) throw Panic("groth16 prove failed: " + pd.err, "{}");   // panic(err), provers.go:148-151,216-219
    return output_json(pd, pb);
}

// ---- ProveBatch: a JSON array of InputParams -> a JSON array of OutputParams / error payloads, in order.
// Every cipher's share goes to the GPUs as ONE g16_prove_*_batch call on the cipher's multi-device handle (request i of the
// share -> device i mod G).
std::string prove_batch_impl(const uint8_t* params, size_t len) {
    Json j{(const char*)params, (const char*)params, (const char*)params + len};
    j.ws();
    if (j.p < j.end && *j.p != '[')
        throw Panic(std::string("json: cannot unmarshal ") + j.kind_name() + " into Go value of type []*impl.InputParams", "{}");
    j.expect('[');
    std::vector<InputParams> reqs;
    std::vector<std::string> results;   // filled early for requests that fail validation
    std::vector<int> algs;
    if (!j.eat(']')) {
        do {
            reqs.push_back(parse_params_at(j));   // a syntax error anywhere fails the whole call, as one Unmarshal would
        } while (j.eat(','));
        j.expect(']');
    }
    j.ws();
    if (j.p < j.end) syntax_error(std::string("invalid character '") + *j.p + "' after top-level value", j.off() + 1);
    const size_t n = reqs.size();
    results.assign(n, std::string());
    algs.assign(n, -1);
    for (size_t i = 0; i < n; i++) {
        try {
            algs[i] = validate_request(reqs[i]);
        } catch (const std::exception& e) {
            results[i] = panic_payload(e);
        }
    }
    for (int alg = 0; alg < 3; alg++) {
        std::vector<size_t> idx;
        for (size_t i = 0; i < n; i++) if (algs[i] == alg) idx.push_back(i);
        if (idx.empty()) continue;
        g16_ctx* ctx;
        {
            std::lock_guard<std::mutex> lk(g_mu);
            ctx = g_provers[alg];
        }
        const size_t m = idx.size(), pb = alg == 0 ? 164 : 196, klen = reqs[idx[0]].key.size();
        std::vector<uint8_t> keys(klen * m), nonces(12 * m), inputs(64 * m), proofs(pb * m), cts(64 * m);
        std::vector<uint32_t> counters(m), st(m, 0);
        for (size_t k = 0; k < m; k++) {
            const InputParams& ip = reqs[idx[k]];
            memcpy(&keys[klen * k], ip.key.data(), klen);
            memcpy(&nonces[12 * k], ip.nonce.data(), 12);
            memcpy(&inputs[64 * k], ip.input.data(), 64);
            counters[k] = ip.counter;
        }
        int rc = alg == 0 ? g16_prove_chacha_batch(ctx, m, keys.data(), nonces.data(), counters.data(), inputs.data(), nullptr, proofs.data(), cts.data())
                          : g16_prove_aes_batch(ctx, m, keys.data(), klen, nonces.data(), counters.data(), inputs.data(), nullptr, proofs.data(), cts.data());
        const std::string msg = rc ? g16_last_error() : "";
        if (rc == G16_ERR_UNSAT && g16_last_batch_status(ctx, st.data(), m) != 0) st.assign(m, 1);
        for (size_t k = 0; k < m; k++) {
            const bool bad = rc == G16_ERR_UNSAT ? st[k] != 0 : rc != 0;
            if (bad) { results[idx[k]] = "{}"; printf("groth16 prove failed: %s\n", msg.c_str()); continue; }
            Pending pd;
            memcpy(pd.proof, &proofs[pb * k], pb);
            memcpy(pd.ct, &cts[64 * k], 64);
            results[idx[k]] = output_json(pd, pb);
        }
        volatile uint8_t* kp = keys.data();
        for (size_t i = 0; i < keys.size(); i++) kp[i] = 0;
    }
    std::string out = "[";
    for (size_t i = 0; i < n; i++) { if (i) out += ","; out += results[i]; }
    return out + "]";
}

// G16_DEVICES = "all" | "0,1,2,..." ; default: G16_DEVICE (one index), else device 0
std::vector<int> device_list() {
    std::vector<int> devs;
    const char* s = getenv("G16_DEVICES");
    if (s && *s) {
        if (!strcmp(s, "all")) {
            int n = 0;
            if (g16_device_count(&n) == 0) for (int i = 0; i < n; i++) devs.push_back(i);
        } else {
            const char* p = s;
            while (*p) {
                char* e = nullptr;
                long v = strtol(p, &e, 10);
                if (e == p) break;
                devs.push_back((int)v);
                p = *e == ',' ? e + 1 : e;
                if (*e && *e != ',') break;
            }
        }
    }
    if (devs.empty()) {
        int device = 0;
        if (const char* d = getenv("G16_DEVICE")) device = atoi(d);
        devs.push_back(device);
    }
    return devs;
}

// ---- libraries/verifier: InputVerifyParams (verify_impl.go:18-22) and the public-witness layouts of verifiers.go:50-152
struct VerifyParams {
    std::string cipher;
    std::vector<uint8_t> proof, signals;
};
VerifyParams parse_verify_params(const uint8_t* data, size_t len) {
    Json j{(const char*)data, (const char*)data, (const char*)data + len};
    VerifyParams vp;
    j.expect('{');
    if (!j.eat('}')) {
        do {
            std::string k = j.str();
            j.expect(':');
            if (key_is(k, "cipher")) vp.cipher = j.str();
            else if (key_is(k, "proof")) vp.proof = read_bytes(j, "proof");
            else if (key_is(k, "publicSignals")) vp.signals = read_bytes(j, "publicSignals");
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
    const int device = device_list()[0];
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
    const std::vector<int> devices = device_list();
    g16_ctx* ctx = nullptr;
    int rc = g16_init_multi((const uint8_t*)provingKey.data, (size_t)provingKey.len, (const uint8_t*)r1cs.data, (size_t)r1cs.len,
                            devices.data(), devices.size(), &ctx);
    if (rc) {
        printf("error reading proving key: %s\n", g16_last_error());   // prove_impl.go:88-91,104-107
        return 0;
    }
    g_provers[algorithmID] = ctx;
    const char* dyn = getenv("G16_DYNAMIC_BATCH");
    if (!dyn || atoi(dyn) != 0) {
        Batcher* b = new Batcher();
        b->alg = algorithmID;
        for (size_t k = 0; k < devices.size(); k++) {
            g16_ctx* h = nullptr;
            if (g16_ctx_device_handle(ctx, k, &h) == 0) b->devs.push_back(h);
        }
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
                // every device proves a full-size batch (the multi-device handle shards request i -> device i mod G)
                n *= b->devs.size();
                std::vector<uint8_t> keys(n * 32), nonces(n * 12), inputs(n * 64), proofs(n * 164), cts(n * 64);
                std::vector<uint32_t> counters(n);
                // different requests, not copies of one: the context learns from its first witnesses which wires are bits
                // (combination tables of the wire-driven queries), and a thousand copies of one witness are one sample
                uint64_t x = 0x9E3779B97F4A7C15ull;
                auto next = [&x] { x += 0x9E3779B97F4A7C15ull; uint64_t z = x; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                                   z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); };
                for (auto& b : keys) b = (uint8_t)next();
                for (auto& b : nonces) b = (uint8_t)next();
                for (auto& b : inputs) b = (uint8_t)next();
                for (auto& c : counters) c = (uint32_t)next();
                // twice: the first batch is proved on the general path while the wires are classified, the second one builds the
                // combination tables and runs through them, so the first caller meets neither the build nor a cold kernel
                for (int pass = 0; pass < 2; pass++)
                    if (g16_prove_chacha_batch(ctx, n, keys.data(), nonces.data(), counters.data(), inputs.data(), nullptr,
                                               proofs.data(), cts.data())) {
                        printf("warm-up batch failed: %s\n", g16_last_error());
                        break;
                    }
            }
        }
#endif
        for (size_t k = 0; k < b->devs.size(); k++) b->workers.emplace_back([b, k] { b->run(k); });
        g_batchers[algorithmID] = b;
    }
    return 1;
}

void Free(void* pointer) { free(pointer); }

static Prove_return_g16 to_c_bytes(const std::string& res) {   // C.CBytes, libprove.go:40,46
    Prove_return_g16 r;
    r.r1 = (long long)res.size();
    r.r0 = malloc(res.size() ? res.size() : 1);
    if (r.r0) memcpy(r.r0, res.data(), res.size());
    else r.r1 = 0;
    return r;
}

Prove_return_g16 Prove(GoSlice_g16 params) {
    std::string res;
    try {
        if (!params.data || params.len <= 0) throw Panic("unexpected end of JSON input", "{\"Offset\":0}");
        res = prove_impl((const uint8_t*)params.data, (size_t)params.len);
    } catch (const std::exception& e) {
        printf("%s\n", e.what());          // libprove.go:35
        res = panic_payload(e);            // libprove.go:36-41: json.Marshal(err) returned as the payload
    }
    return to_c_bytes(res);
}

// the batched twin of Prove (Prove itself is unchanged)
Prove_return_g16 ProveBatch(GoSlice_g16 params) {
    std::string res;
    try {
        if (!params.data || params.len <= 0) throw Panic("unexpected end of JSON input", "{\"Offset\":0}");
        res = prove_batch_impl((const uint8_t*)params.data, (size_t)params.len);
    } catch (const std::exception& e) {
        printf("%s\n", e.what());
        res = panic_payload(e);
    }
    return to_c_bytes(res);
}

// statistics of the Prove batcher of one cipher: out[0] batches run, out[1] requests proved, out[2] devices
int g16_libprove_stats(int algorithmID, uint64_t out[3]) {
    if (algorithmID < 0 || algorithmID > 2 || !out) return G16_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_mu);
    Batcher* b = g_batchers[algorithmID];
    if (!b) return G16_ERR_STATE;
    std::lock_guard<std::mutex> lk2(b->mu);
    out[0] = b->batches; out[1] = b->proved; out[2] = b->devs.size();
    return G16_OK;
}

// test hook: drop the cached provers (the reference has no such call; its map lives for the process lifetime)
void g16_libprove_reset(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& b : g_batchers) {
        if (!b) continue;
        { std::lock_guard<std::mutex> lk2(b->mu); b->stop = true; }
        b->cv_work.notify_all();
        for (auto& w : b->workers) if (w.joinable()) w.join();
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
