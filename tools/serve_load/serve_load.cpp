// Load generator for the reference's ONE-REQUEST API (libraries/prover/libprove.go: InitAlgorithm / Prove / Free), in C++ so that
// the callers are not throttled by an interpreter lock: T threads each issue K Prove(JSON) calls back to back against
// libg16b200.so, which coalesces concurrent calls into GPU batches (one batching worker per GPU of G16_DEVICES).
//   serve_load <libg16b200.so> <pk.chacha20> <r1cs.chacha20> <threads> <calls per thread>
// Prints one JSON line: proofs/s, latency percentiles, batcher statistics. Every response is checked to be an OutputParams
// object; the ciphertext of every 64th response is checked against a host ChaCha20 (RFC 7539).
#include <dlfcn.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <system_error>
#include <thread>
#include <vector>

struct GoSlice { void* data; long long len; long long cap; };
struct ProveRet { void* r0; long long r1; };
typedef unsigned char (*InitFn)(unsigned char, GoSlice, GoSlice);
typedef ProveRet (*ProveFn)(GoSlice);
typedef void (*FreeFn)(void*);
typedef int (*StatsFn)(int, uint64_t*);

static std::vector<uint8_t> slurp(const char* p) {
    std::ifstream f(p, std::ios::binary);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static uint32_t rotl(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
static void chacha_block(const uint8_t key[32], uint32_t counter, const uint8_t nonce[12], uint8_t out[64]) {
    uint32_t s[16] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
    for (int i = 0; i < 8; i++) memcpy(&s[4 + i], key + 4 * i, 4);
    s[12] = counter;
    for (int i = 0; i < 3; i++) memcpy(&s[13 + i], nonce + 4 * i, 4);
    uint32_t x[16];
    memcpy(x, s, 64);
#define QR(a, b, c, d) x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12); x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8); x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
    for (int r = 0; r < 10; r++) { QR(0, 4, 8, 12) QR(1, 5, 9, 13) QR(2, 6, 10, 14) QR(3, 7, 11, 15) QR(0, 5, 10, 15) QR(1, 6, 11, 12) QR(2, 7, 8, 13) QR(3, 4, 9, 14) }
    for (int i = 0; i < 16; i++) { x[i] += s[i]; memcpy(out + 4 * i, &x[i], 4); }
}
static int b64v(char c) { return c >= 'A' && c <= 'Z' ? c - 'A' : c >= 'a' && c <= 'z' ? c - 'a' + 26 : c >= '0' && c <= '9' ? c - '0' + 52 : c == '+' ? 62 : c == '/' ? 63 : -1; }

int main(int argc, char** argv) {
    if (argc < 6) { fprintf(stderr, "usage: serve_load lib pk r1cs threads calls\n"); return 2; }
    void* h = dlopen(argv[1], RTLD_NOW);
    if (!h) { fprintf(stderr, "%s\n", dlerror()); return 2; }
    InitFn init = (InitFn)dlsym(h, "InitAlgorithm");
    ProveFn prove = (ProveFn)dlsym(h, "Prove");
    FreeFn freefn = (FreeFn)dlsym(h, "Free");
    StatsFn stats = (StatsFn)dlsym(h, "g16_libprove_stats");
    std::vector<uint8_t> pk = slurp(argv[2]), r1 = slurp(argv[3]);
    const int T = atoi(argv[4]), K = atoi(argv[5]);
    auto t_init = std::chrono::steady_clock::now();
    if (!init(0, GoSlice{pk.data(), (long long)pk.size(), (long long)pk.size()}, GoSlice{r1.data(), (long long)r1.size(), (long long)r1.size()})) {
        fprintf(stderr, "InitAlgorithm failed\n");
        return 1;
    }
    const double init_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_init).count();
    std::vector<std::vector<double>> lat(T);
    std::atomic<int> bad{0};
    std::atomic<long long> done{0};
    auto body = [&](int t) {
        lat[t].reserve(K);
        for (int k = 0; k < K; k++) {
            uint8_t key[32], nonce[12], in[64];
            uint32_t seed = (uint32_t)(t * 2654435761u + k * 40503u + 12345u);
            auto next = [&] { seed = seed * 1664525u + 1013904223u; return (uint8_t)(seed >> 24); };
            for (auto& b : key) b = next();
            for (auto& b : nonce) b = next();
            for (auto& b : in) b = next();
            const uint32_t counter = seed;
            std::string js = "{\"cipher\":\"chacha20\",\"key\":[";
            auto arr = [&](const uint8_t* p, int n) { for (int i = 0; i < n; i++) { js += std::to_string(p[i]); if (i + 1 < n) js += ','; } };
            arr(key, 32); js += "],\"nonce\":["; arr(nonce, 12); js += "],\"counter\":" + std::to_string(counter) + ",\"input\":["; arr(in, 64); js += "]}";
            auto t0 = std::chrono::steady_clock::now();
            ProveRet r = prove(GoSlice{(void*)js.data(), (long long)js.size(), (long long)js.size()});
            lat[t].push_back(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
            std::string out((const char*)r.r0, (size_t)r.r1);
            freefn(r.r0);
            if (out.compare(0, 20, "{\"proof\":{\"proofJson") != 0) { bad++; continue; }
            if (((t + k) & 63) == 0) {   // ciphertext check: publicSignals = base64(ChaCha20(key, nonce, counter) xor input)
                size_t p = out.find("\"publicSignals\":\"");
                uint8_t ks[64], ct[64];
                chacha_block(key, counter, nonce, ks);
                int bits = 0, n = 0;
                uint32_t acc = 0;
                for (size_t i = p + 17; i < out.size() && n < 64; i++) {
                    int v = b64v(out[i]);
                    if (v < 0) break;
                    acc = (acc << 6) | (uint32_t)v; bits += 6;
                    if (bits >= 8) { bits -= 8; ct[n++] = (uint8_t)(acc >> bits); }
                }
                bool okc = n == 64;
                for (int i = 0; i < 64 && okc; i++) okc = ct[i] == (uint8_t)(ks[i] ^ in[i]);
                if (!okc) bad++;
            }
            done++;
        }
    };
    // InitAlgorithm has already proved a full-size warm-up batch on every device (G16_PREWARM)
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    int started = 0;
    for (int t = 0; t < T; t++) {
        try { th.emplace_back(body, t); started++; }
        catch (const std::system_error&) { break; }   // the box's thread limit: run with the callers that could be created
    }
    if (started < T) fprintf(stderr, "serve_load: only %d of %d caller threads could be created\n", started, T);
    for (auto& x : th) x.join();
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::vector<double> all;
    for (auto& v : lat) all.insert(all.end(), v.begin(), v.end());
    std::sort(all.begin(), all.end());
    auto pct = [&](double q) { return all.empty() ? 0.0 : all[(size_t)(q * (all.size() - 1))]; };
    uint64_t st[3] = {0, 0, 0};
    if (stats) stats(0, st);
    printf("{\"callers\": %d, \"calls_per_caller\": %d, \"proofs\": %lld, \"bad\": %d, \"seconds\": %.3f, \"proofs_per_s\": %.1f, "
           "\"latency_ms\": {\"p50\": %.2f, \"p95\": %.2f, \"p99\": %.2f, \"max\": %.2f}, \"init_s\": %.2f, "
           "\"batches\": %llu, \"devices\": %llu, \"mean_batch\": %.1f, \"api\": \"Prove(JSON), one request per call, C++ caller threads\"}\n",
           started, K, (long long)done.load(), bad.load(), s, done.load() / s, pct(0.5), pct(0.95), pct(0.99), pct(1.0), init_s,
           (unsigned long long)st[0], (unsigned long long)st[2], st[0] ? (double)st[1] / (double)st[0] : 0.0);
    return bad.load() ? 1 : 0;
}
