// Device kernels of the Groth16 prove path other than MSM / NTT:
//   * point decompression of the proving key            (replaces gnark marshal.go:311-348 ProvingKey.readFrom — a19)
//   * ChaCha20 witness assignment                        (replaces provers.go:79-142 + utils/bytes.go:11-47 — a3..a6)
//   (the batched R1CS solver lives in solver.cuh / k_solver.cu)
//   (proof assembly + serialisation live in assemble.cuh / k_assemble.cu, a hot translation unit)
#pragma once
#include "fixed_base.cuh"
#include "prover_api.hpp"
#include "serialize.cuh"

namespace g16 {

// ------------------------------------------------------------------------------------------------ decompression
FD Fp fp_three() { Fp o = Fp::one(); return o + o + o; }
FD Fp2 g2_coeff_b() {   // 3/(9+u)
    Fp o = Fp::one();
    Fp three = o + o + o;
    Fp nine = three + three + three;
    Fp2 xi = {nine, o};
    return Fp2{three, Fp::zero()} * xi.inv();
}
// exponents derived from p: which = 0 (p+1)/4, 1 (p-3)/4, 2 (p-1)/2
FD void fp_exponent(int which, uint32_t e[8]) {
    uint32_t t[8];
    for (int i = 0; i < 8; i++) t[i] = FpParams::mod(i);
    int sh;
    if (which == 0) { t[0] += 1; sh = 2; }        // low limb ...47: no carry
    else if (which == 1) { t[0] -= 3; sh = 2; }
    else { t[0] -= 1; sh = 1; }
    for (int i = 0; i < 8; i++) e[i] = (t[i] >> sh) | (i < 7 ? (t[i + 1] << (32 - sh)) : 0u);
}

// one compressed point -> affine Montgomery. Returns 0, or error bits: 1 = not on the curve, 2 = bad flag bits,
// 4 = a coordinate encoded as an integer >= p (SetBytesCanonical), 8 = infinity flag with non-zero payload
// (gnark-crypto ecc/bn254/marshal.go: 0b10 smallest y, 0b11 largest y, 0b01 infinity, 0b00 = uncompressed, not accepted here)
FD uint32_t decompress_g1_point(const uint8_t* b, G1Affine& out) {
    out = G1Affine::inf();
    uint8_t flag = b[0] & 0xC0;
    if (flag == 0x40) {
        uint32_t o = b[0] & 0x3F;
        for (int i = 1; i < 32; i++) o |= b[i];
        return o ? 8u : 0u;
    }
    if (flag == 0x00) return 2;
    bool nc = false;
    Fp x = fp_from_be32(b, true, &nc);
    if (nc) return 4;
    Fp rhs = x.sqr() * x + fp_three();
    uint32_t e[8];
    fp_exponent(0, e);
    Fp y = rhs.pow(e);
    if (y.sqr() != rhs) return 1;
    bool want_largest = flag == 0xC0;
    if (y.lex_largest() != want_largest) y = y.neg();
    out = {x, y};
    return 0;
}
FD uint32_t decompress_g2_point(const uint8_t* b, G2Affine& out) {
    out = G2Affine::inf();
    uint8_t flag = b[0] & 0xC0;
    if (flag == 0x40) {
        uint32_t o = b[0] & 0x3F;
        for (int i = 1; i < 64; i++) o |= b[i];
        return o ? 8u : 0u;
    }
    if (flag == 0x00) return 2;
    bool nc = false;
    Fp2 x = {fp_from_be32(b + 32, false, &nc), fp_from_be32(b, true, &nc)};   // X.A1 || X.A0
    if (nc) return 4;
    Fp2 a = x.sqr() * x + g2_coeff_b();
    Fp2 y;
    if (a.is_zero()) {
        y = a;
    } else {
        // square root in Fp2, p = 3 mod 4 (Adj & Rodriguez-Henriquez, alg. 9)
        uint32_t e[8];
        fp_exponent(1, e);
        Fp2 a1 = a.pow(e);
        Fp2 x0 = a1 * a;
        Fp2 alpha = a1 * x0;
        if (alpha == Fp2::one().neg()) {
            y = Fp2{Fp::zero(), Fp::one()} * x0;
        } else {
            fp_exponent(2, e);
            Fp2 bb = (Fp2::one() + alpha).pow(e);
            y = bb * x0;
        }
        if (y.sqr() != a) return 1;
    }
    bool want_largest = flag == 0xC0;
    if (y.lex_largest() != want_largest) y = y.neg();
    out = {x, y};
    return 0;
}
// in: 32-byte compressed G1 points; out: affine Montgomery. err: bit 0 = not on curve, bit 1 = bad flag
__global__ void decompress_g1_kernel(const uint8_t* __restrict__ in, uint32_t n, G1Affine* __restrict__ out,
                                     uint32_t* __restrict__ err) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Affine p;
    uint32_t e = decompress_g1_point(in + (size_t)i * 32, p);
    if (e) atomicOr(err, e);
    out[i] = p;
}
__global__ void decompress_g2_kernel(const uint8_t* __restrict__ in, uint32_t n, G2Affine* __restrict__ out,
                                     uint32_t* __restrict__ err) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G2Affine p;
    uint32_t e = decompress_g2_point(in + (size_t)i * 64, p);
    if (e) atomicOr(err, e);
    out[i] = p;
}

// big-endian 32-byte canonical scalars -> Fr canonical limbs (NOT Montgomery)
__global__ void scalars_from_be_kernel(const uint8_t* __restrict__ in, uint32_t n, Fr* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* b = in + (size_t)i * 32;
    Fr v;
    for (int k = 0; k < 8; k++) {
        const uint8_t* q = b + 28 - 4 * k;
        v.l[k] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
    // reduce into [0, r): inputs are < 2^256 < 6r
    for (int k = 0; k < 5; k++) v.reduce_once();
    out[i] = v;
}

// ------------------------------------------------------------------------------------------------ ChaCha20 witness
FD uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
#define G16_QR(a, b, c, d)                                   \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16);            \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12);            \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);             \
    x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);

FD uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
FD uint32_t be32(const uint8_t* p) { return (uint32_t)p[3] | ((uint32_t)p[2] << 8) | ((uint32_t)p[1] << 16) | ((uint32_t)p[0] << 24); }

// One thread per request: ciphertext = ChaCha20(key, nonce, counter) xor input (RFC 7539 block function; the reference
// calls x/crypto/chacha20, provers.go:95-101), then the witness in gnark's order: ONE | Counter[32] | Nonce[3][32] |
// In[16][32] | Out[16][32] | Key[8][32], every word LSB-first; In/Out words are loaded big-endian, Key/Nonce words
// little-endian (provers.go:106-142, utils/bytes.go:11-47).
__global__ void chacha_witness_kernel(const uint8_t* __restrict__ keys, const uint8_t* __restrict__ nonces,
                                      const uint32_t* __restrict__ counters, const uint8_t* __restrict__ inputs,
                                      uint32_t n, Fr* __restrict__ W, size_t w_stride, uint8_t* __restrict__ ct_out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* key = keys + (size_t)i * 32;
    const uint8_t* nonce = nonces + (size_t)i * 12;
    const uint8_t* in = inputs + (size_t)i * 64;
    uint32_t st[16], x[16];
    st[0] = 0x61707865u; st[1] = 0x3320646eu; st[2] = 0x79622d32u; st[3] = 0x6b206574u;
    for (int k = 0; k < 8; k++) st[4 + k] = le32(key + 4 * k);
    st[12] = counters[i];
    for (int k = 0; k < 3; k++) st[13 + k] = le32(nonce + 4 * k);
    for (int k = 0; k < 16; k++) x[k] = st[k];
    for (int r = 0; r < 10; r++) {
        G16_QR(0, 4, 8, 12) G16_QR(1, 5, 9, 13) G16_QR(2, 6, 10, 14) G16_QR(3, 7, 11, 15)
        G16_QR(0, 5, 10, 15) G16_QR(1, 6, 11, 12) G16_QR(2, 7, 8, 13) G16_QR(3, 4, 9, 14)
    }
    uint8_t ct[64];
    for (int k = 0; k < 16; k++) {
        uint32_t ks = x[k] + st[k];
        for (int b = 0; b < 4; b++) ct[4 * k + b] = in[4 * k + b] ^ (uint8_t)(ks >> (8 * b));
    }
    for (int k = 0; k < 64; k++) ct_out[(size_t)i * 64 + k] = ct[k];
    // wire-major layout: wire k of request i lives at W[k * w_stride + i] (coalesced across the requests of a warp)
    Fr* w = W + i;
    const Fr one = Fr::one(), zero = Fr::zero();
    size_t pos = 0;
    w[pos] = one;
    pos += w_stride;
    auto put_word = [&](uint32_t word) {
        for (int b = 0; b < 32; b++) { w[pos] = ((word >> b) & 1u) ? one : zero; pos += w_stride; }
    };
    put_word(counters[i]);
    for (int k = 0; k < 3; k++) put_word(le32(nonce + 4 * k));
    for (int k = 0; k < 16; k++) put_word(be32(in + 4 * k));
    for (int k = 0; k < 16; k++) put_word(be32(ct + 4 * k));
    for (int k = 0; k < 8; k++) put_word(le32(key + 4 * k));
}

// generic: wire 0 = 1, wires 1..nw = witness[i][..]   (wire-major output, see above)
__global__ void witness_copy_kernel(const Fr* __restrict__ witness, uint32_t n_witness, uint32_t batch, Fr* __restrict__ W,
                                    size_t w_stride) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)batch * (n_witness + 1)) return;
    uint32_t i = (uint32_t)(gid / (n_witness + 1)), k = (uint32_t)(gid % (n_witness + 1));
    W[(size_t)k * w_stride + i] = k == 0 ? Fr::one() : witness[(size_t)i * n_witness + (k - 1)];
}

// ------------------------------------------------------------------------------------------------ AES-CTR witness
// FIPS-197 S-box. (The reference computes the ciphertext with crypto/aes + cipher.NewCTR, provers.go:184-192.)
__device__ const uint8_t G16_AES_SBOX[256] = {
    0x63,0x7c,0x77,0x7b,0xf2,0x6b,0x6f,0xc5,0x30,0x01,0x67,0x2b,0xfe,0xd7,0xab,0x76,0xca,0x82,0xc9,0x7d,0xfa,0x59,0x47,0xf0,
    0xad,0xd4,0xa2,0xaf,0x9c,0xa4,0x72,0xc0,0xb7,0xfd,0x93,0x26,0x36,0x3f,0xf7,0xcc,0x34,0xa5,0xe5,0xf1,0x71,0xd8,0x31,0x15,
    0x04,0xc7,0x23,0xc3,0x18,0x96,0x05,0x9a,0x07,0x12,0x80,0xe2,0xeb,0x27,0xb2,0x75,0x09,0x83,0x2c,0x1a,0x1b,0x6e,0x5a,0xa0,
    0x52,0x3b,0xd6,0xb3,0x29,0xe3,0x2f,0x84,0x53,0xd1,0x00,0xed,0x20,0xfc,0xb1,0x5b,0x6a,0xcb,0xbe,0x39,0x4a,0x4c,0x58,0xcf,
    0xd0,0xef,0xaa,0xfb,0x43,0x4d,0x33,0x85,0x45,0xf9,0x02,0x7f,0x50,0x3c,0x9f,0xa8,0x51,0xa3,0x40,0x8f,0x92,0x9d,0x38,0xf5,
    0xbc,0xb6,0xda,0x21,0x10,0xff,0xf3,0xd2,0xcd,0x0c,0x13,0xec,0x5f,0x97,0x44,0x17,0xc4,0xa7,0x7e,0x3d,0x64,0x5d,0x19,0x73,
    0x60,0x81,0x4f,0xdc,0x22,0x2a,0x90,0x88,0x46,0xee,0xb8,0x14,0xde,0x5e,0x0b,0xdb,0xe0,0x32,0x3a,0x0a,0x49,0x06,0x24,0x5c,
    0xc2,0xd3,0xac,0x62,0x91,0x95,0xe4,0x79,0xe7,0xc8,0x37,0x6d,0x8d,0xd5,0x4e,0xa9,0x6c,0x56,0xf4,0xea,0x65,0x7a,0xae,0x08,
    0xba,0x78,0x25,0x2e,0x1c,0xa6,0xb4,0xc6,0xe8,0xdd,0x74,0x1f,0x4b,0xbd,0x8b,0x8a,0x70,0x3e,0xb5,0x66,0x48,0x03,0xf6,0x0e,
    0x61,0x35,0x57,0xb9,0x86,0xc1,0x1d,0x9e,0xe1,0xf8,0x98,0x11,0x69,0xd9,0x8e,0x94,0x9b,0x1e,0x87,0xe9,0xce,0x55,0x28,0xdf,
    0x8c,0xa1,0x89,0x0d,0xbf,0xe6,0x42,0x68,0x41,0x99,0x2d,0x0f,0xb0,0x54,0xbb,0x16};

FD uint8_t aes_xtime(uint8_t x) { return (uint8_t)((x << 1) ^ ((x >> 7) * 0x1b)); }

// expanded key: 4*(Nr+1) words as bytes; Nk = 4 (AES-128, Nr = 10) or 8 (AES-256, Nr = 14)
__device__ __forceinline__ void aes_expand_key(const uint8_t* key, int nk, uint8_t* rk /* 240 bytes */) {
    int nr = nk + 6;
    for (int i = 0; i < 4 * nk; i++) rk[i] = key[i];
    uint8_t rcon = 1;
    for (int i = nk; i < 4 * (nr + 1); i++) {
        uint8_t t[4] = {rk[4 * (i - 1)], rk[4 * (i - 1) + 1], rk[4 * (i - 1) + 2], rk[4 * (i - 1) + 3]};
        if (i % nk == 0) {
            uint8_t u = t[0];
            t[0] = G16_AES_SBOX[t[1]] ^ rcon; t[1] = G16_AES_SBOX[t[2]]; t[2] = G16_AES_SBOX[t[3]]; t[3] = G16_AES_SBOX[u];
            rcon = aes_xtime(rcon);
        } else if (nk > 6 && i % nk == 4) {
            for (int k = 0; k < 4; k++) t[k] = G16_AES_SBOX[t[k]];
        }
        for (int k = 0; k < 4; k++) rk[4 * i + k] = rk[4 * (i - nk) + k] ^ t[k];
    }
}
__device__ __forceinline__ void aes_encrypt_block(const uint8_t* rk, int nr, const uint8_t in[16], uint8_t out[16]) {
    uint8_t s[16], t[16];
    for (int i = 0; i < 16; i++) s[i] = in[i] ^ rk[i];
    for (int r = 1; r <= nr; r++) {
        // SubBytes + ShiftRows (state is column-major: s[4c + r])
        for (int c = 0; c < 4; c++)
            for (int rw = 0; rw < 4; rw++) t[4 * c + rw] = G16_AES_SBOX[s[4 * ((c + rw) & 3) + rw]];
        if (r < nr) {
            for (int c = 0; c < 4; c++) {
                uint8_t a0 = t[4 * c], a1 = t[4 * c + 1], a2 = t[4 * c + 2], a3 = t[4 * c + 3];
                uint8_t x = a0 ^ a1 ^ a2 ^ a3;
                s[4 * c] = a0 ^ x ^ aes_xtime(a0 ^ a1);
                s[4 * c + 1] = a1 ^ x ^ aes_xtime(a1 ^ a2);
                s[4 * c + 2] = a2 ^ x ^ aes_xtime(a2 ^ a3);
                s[4 * c + 3] = a3 ^ x ^ aes_xtime(a3 ^ a0);
            }
        } else {
            for (int i = 0; i < 16; i++) s[i] = t[i];
        }
        for (int i = 0; i < 16; i++) s[i] ^= rk[16 * r + i];
    }
    for (int i = 0; i < 16; i++) out[i] = s[i];
}
FD Fr fr_from_u32(uint32_t v) {
    Fr x = Fr::zero();
    x.l[0] = v;
    return x.to_mont();
}
// One thread per request. CTR block = nonce(12) || BE32(counter + block) (provers.go:191, aesV2/common.go:113-120),
// 4 blocks. Witness order: ONE | Nonce[12] | Counter | Plaintext[64] | Ciphertext[64] | Key[key_len], one byte per wire.
__global__ void aes_witness_kernel(const uint8_t* __restrict__ keys, uint32_t key_len, const uint8_t* __restrict__ nonces,
                                   const uint32_t* __restrict__ counters, const uint8_t* __restrict__ inputs, uint32_t n,
                                   Fr* __restrict__ W, size_t w_stride, uint8_t* __restrict__ ct_out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* key = keys + (size_t)i * key_len;
    const uint8_t* nonce = nonces + (size_t)i * 12;
    const uint8_t* in = inputs + (size_t)i * 64;
    uint8_t rk[240];
    int nk = (int)key_len / 4;
    aes_expand_key(key, nk, rk);
    uint8_t ct[64];
    uint32_t ctr = counters[i];
    // cipher.NewCTR increments the whole 16-byte block as one big-endian integer: a counter within 3 of 2^32 carries into
    // the nonce bytes. (The circuit then rejects the witness, aes128.go:49-50 AssertIsLessOrEqual(counter, MaxUint32),
    // exactly as groth16.Prove fails in the reference.)
    uint8_t blk[16];
    for (int k = 0; k < 12; k++) blk[k] = nonce[k];
    blk[12] = (uint8_t)(ctr >> 24); blk[13] = (uint8_t)(ctr >> 16); blk[14] = (uint8_t)(ctr >> 8); blk[15] = (uint8_t)ctr;
    for (int b = 0; b < 4; b++) {
        uint8_t ks[16];
        aes_encrypt_block(rk, nk + 6, blk, ks);
        for (int k = 0; k < 16; k++) ct[16 * b + k] = in[16 * b + k] ^ ks[k];
        for (int k = 15; k >= 0; k--) if (++blk[k] != 0) break;
    }
    for (int k = 0; k < 64; k++) ct_out[(size_t)i * 64 + k] = ct[k];
    Fr* w = W + i;
    size_t pos = 0;
    w[pos] = Fr::one(); pos += w_stride;
    for (int k = 0; k < 12; k++) { w[pos] = fr_from_u32(nonce[k]); pos += w_stride; }
    w[pos] = fr_from_u32(ctr); pos += w_stride;
    for (int k = 0; k < 64; k++) { w[pos] = fr_from_u32(in[k]); pos += w_stride; }
    for (int k = 0; k < 64; k++) { w[pos] = fr_from_u32(ct[k]); pos += w_stride; }
    for (uint32_t k = 0; k < key_len; k++) { w[pos] = fr_from_u32(key[k]); pos += w_stride; }
}

// ------------------------------------------------------------------------------------------------ BSB22 commitment hash
// SHA-256 of a short message held in thread-local memory (<= 183 bytes -> at most 3 blocks)
__device__ __forceinline__ uint32_t sha_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
__device__ const uint32_t G16_SHA256_K[64] = {
    0x428a2f98,0x71374491,0xb5c0fbcf,0xe9b5dba5,0x3956c25b,0x59f111f1,0x923f82a4,0xab1c5ed5,0xd807aa98,0x12835b01,0x243185be,0x550c7dc3,
    0x72be5d74,0x80deb1fe,0x9bdc06a7,0xc19bf174,0xe49b69c1,0xefbe4786,0x0fc19dc6,0x240ca1cc,0x2de92c6f,0x4a7484aa,0x5cb0a9dc,0x76f988da,
    0x983e5152,0xa831c66d,0xb00327c8,0xbf597fc7,0xc6e00bf3,0xd5a79147,0x06ca6351,0x14292967,0x27b70a85,0x2e1b2138,0x4d2c6dfc,0x53380d13,
    0x650a7354,0x766a0abb,0x81c2c92e,0x92722c85,0xa2bfe8a1,0xa81a664b,0xc24b8b70,0xc76c51a3,0xd192e819,0xd6990624,0xf40e3585,0x106aa070,
    0x19a4c116,0x1e376c08,0x2748774c,0x34b0bcb5,0x391c0cb3,0x4ed8aa4a,0x5b9cca4f,0x682e6ff3,0x748f82ee,0x78a5636f,0x84c87814,0x8cc70208,
    0x90befffa,0xa4506ceb,0xbef9a3f7,0xc67178f2};
__device__ __forceinline__ void sha256_short(const uint8_t* msg, uint32_t len, uint8_t out[32]) {
    uint8_t buf[192];
    uint32_t total = ((len + 9 + 63) / 64) * 64;
    for (uint32_t i = 0; i < total; i++) buf[i] = i < len ? msg[i] : 0;
    buf[len] = 0x80;
    uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; i++) buf[total - 1 - i] = (uint8_t)(bits >> (8 * i));
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    for (uint32_t blk = 0; blk < total; blk += 64) {
        uint32_t w[64];
        for (int i = 0; i < 16; i++)
            w[i] = ((uint32_t)buf[blk + 4 * i] << 24) | ((uint32_t)buf[blk + 4 * i + 1] << 16) | ((uint32_t)buf[blk + 4 * i + 2] << 8) | buf[blk + 4 * i + 3];
        for (int i = 16; i < 64; i++) {
            uint32_t s0 = sha_rotr(w[i - 15], 7) ^ sha_rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
            uint32_t s1 = sha_rotr(w[i - 2], 17) ^ sha_rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; i++) {
            uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
            uint32_t ch = (e & f) ^ (~e & g);
            uint32_t t1 = hh + S1 + ch + G16_SHA256_K[i] + w[i];
            uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
            uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            uint32_t t2 = S0 + mj;
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    for (int i = 0; i < 8; i++) { out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16); out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i]; }
}
// challenge = hash_to_field(C.Marshal(), DST "bsb22-commitment"): RFC 9380 expand_message_xmd(SHA-256) to 48 bytes, reduced
// mod r (gnark-crypto fr.Hash; SURVEY.md Appendix F.3 — recalled, self-consistent with oracle/setup.py). One thread per proof.
__global__ void bsb22_challenge_kernel(const G1XYZZ* __restrict__ commit, uint32_t n, Fr* __restrict__ W, size_t w_stride,
                                       uint32_t commit_wire, G1Affine* __restrict__ commit_aff) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1Affine c = commit[i].to_affine();
    commit_aff[i] = c;
    const uint8_t dst_prime[17] = {'b', 's', 'b', '2', '2', '-', 'c', 'o', 'm', 'm', 'i', 't', 'm', 'e', 'n', 't', 16};
    uint8_t m[148];
    for (int k = 0; k < 64; k++) m[k] = 0;                         // Z_pad
    if (c.is_inf()) { for (int k = 0; k < 64; k++) m[64 + k] = 0; m[64] = 0x40; }
    else { fp_to_be32(c.x, m + 64); fp_to_be32(c.y, m + 96); }     // uncompressed X || Y
    m[128] = 0; m[129] = 48; m[130] = 0;                           // l_i_b_str(48), I2OSP(0,1)
    for (int k = 0; k < 17; k++) m[131 + k] = dst_prime[k];
    uint8_t b0[32], b1[32], b2[32], t[50];
    sha256_short(m, 148, b0);
    for (int k = 0; k < 32; k++) t[k] = b0[k];
    t[32] = 1;
    for (int k = 0; k < 17; k++) t[33 + k] = dst_prime[k];
    sha256_short(t, 50, b1);
    for (int k = 0; k < 32; k++) t[k] = b0[k] ^ b1[k];
    t[32] = 2;
    sha256_short(t, 50, b2);
    // 48 big-endian bytes mod r by Horner on canonical values (add / double are representation-agnostic)
    Fr acc = Fr::zero();
    for (int k = 0; k < 48; k++) {
        uint8_t byte = k < 32 ? b1[k] : b2[k - 32];
        for (int d = 0; d < 8; d++) acc = acc.dbl();
        Fr bv = Fr::zero();
        bv.l[0] = byte;
        acc = acc + bv;
    }
    W[(size_t)commit_wire * w_stride + i] = acc.to_mont();
}
// stage-level test entry (g16_bsb22_challenge): affine points -> the XYZZ form bsb22_challenge_kernel consumes
__global__ void g1_affine_to_xyzz_kernel(const G1Affine* __restrict__ in, uint32_t n, G1XYZZ* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = G1XYZZ::from_affine(in[i]);
}
__global__ void assemble_commitment_kernel(const G1Affine* __restrict__ commit_aff, const G1XYZZ* __restrict__ pok, uint32_t n,
                                           uint8_t* __restrict__ out, size_t out_stride) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t* o = out + (size_t)i * out_stride;
    o[128] = 0; o[129] = 0; o[130] = 0; o[131] = 1;
    g1_compress(commit_aff[i], o + 132);
    g1_compress(pok[i].to_affine(), o + 164);
}

// G2 proof element: Bs = msmB2 + beta2 + s*delta2, compressed straight into bytes [32, 96) of the proof (Appendix C, F.1).
// grid n, block 64 (the fixed-base product is a depth-6 tree, see fixed_base.cuh). Runs on the side stream, right behind
// the G2 MSM it consumes, so it is off the critical path of the G1 assembly.
__global__ void __launch_bounds__(FB_WINDOWS)
assemble_g2_kernel(AssemblyKeys keys, uint32_t n, const G2XYZZ* __restrict__ mB2, const Fr* __restrict__ rs,
                   uint8_t* __restrict__ out, size_t out_stride) {
    __shared__ G2XYZZ sm[FB_WINDOWS];
    const uint32_t i = blockIdx.x;
    G2XYZZ v = fixed_base_mul_block<G2>(keys.delta2_tab, rs[2 * i + 1], sm);
    if (threadIdx.x) return;
    v.add(mB2[i]);
    v.madd(keys.beta2, false);
    g2_compress(v.to_affine(), out + (size_t)i * out_stride + 32);
}

// wire-major W -> one row of nb_wires values per witness (test / gnark-shaped output only)
__global__ void wires_to_rows_kernel(const Fr* __restrict__ W, size_t w_stride, uint32_t batch, uint32_t nb_wires,
                                     Fr* __restrict__ out) {
    size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size_t)batch * nb_wires) return;
    uint32_t i = (uint32_t)(gid % batch), k = (uint32_t)(gid / batch);
    out[(size_t)i * nb_wires + k] = W[(size_t)k * w_stride + i];
}

// ------------------------------------------------------------------------------------------------ groth16.Verify
// (gnark v0.11.0 backend/groth16/bn254/verify.go, reached from libraries/verifier/impl/verifiers.go:93-99,139-145; SURVEY
// Appendix C for the proof bytes, F.4 for the equation.)
// Per proof i the pairing inputs are  P[4i..] = (-Ar, alpha, kSum, Krs),  Q[4i..] = (Bs, beta2, gamma2, delta2), and with one
// BSB22 commitment  P2[2i..] = (C, PoK),  Q2[2i..] = (G, GRootSigmaNeg).
// role (blockIdx.y) 0: Ar + the fixed entries + the commitment count of the proof trailer ; 1: Krs ; 2: Bs ; 3: C ; 4: PoK.
// bad[i] != 0 marks a malformed proof (flag bits, point not on its curve, wrong commitment count): it is rejected whatever
// the pairing says. Bs is checked to be on the twist AND in the r-torsion subgroup, as gnark's decoder does (G1 has cofactor 1).
// r-torsion test for a point of the twist, as gnark-crypto v0.14.0 ecc/bn254/g2.go IsInSubGroup (which G2Affine.SetBytes runs
// on every decoded point): [r]P = 0  <=>  [x0+1]P + psi([x0]P) + psi^2([x0]P) = psi^3([2 x0]P), x0 = 4965661367192848881 the
// curve seed, psi = untwist-Frobenius-twist: psi(x, y) = (conj(x) gx, conj(y) gy), gx = xi^((p-1)/3), gy = xi^((p-1)/2).
FD G2XYZZ g2_psi(const G2XYZZ& p, const Fp2& gx, const Fp2& gy) {
    return {p.X.conj() * gx, p.Y.conj() * gy, p.ZZ.conj(), p.ZZZ.conj()};
}
FD bool g2_xyzz_equal(const G2XYZZ& a, const G2XYZZ& b) {
    if (a.is_inf() || b.is_inf()) return a.is_inf() && b.is_inf();
    return a.X * b.ZZ == b.X * a.ZZ && a.Y * b.ZZZ == b.Y * a.ZZZ;
}
FD bool g2_in_subgroup(const G2Affine& q, const Fp2& gx, const Fp2& gy) {
    if (q.is_inf()) return true;
    Scalar256 x0;
    x0.w[0] = 0x4a6909f1u; x0.w[1] = 0x44e992b4u;   // 4965661367192848881
    for (int i = 2; i < 8; i++) x0.w[i] = 0;
    const G2XYZZ P = G2XYZZ::from_affine(q);
    const G2XYZZ a = scalar_mul(P, x0);          // [x0]P
    const G2XYZZ b = g2_psi(a, gx, gy);          // psi([x0]P)
    G2XYZZ lhs = a;
    lhs.add(P);                                  // [x0+1]P
    lhs.add(b);
    lhs.add(g2_psi(b, gx, gy));                  // + psi^2([x0]P)
    const G2XYZZ rhs = g2_psi(g2_psi(g2_psi(a.dbl(), gx, gy), gx, gy), gx, gy);
    return g2_xyzz_equal(lhs, rhs);
}
__global__ void g2_subgroup_kernel(const G2Affine* __restrict__ pts, uint32_t n, const Fp2* __restrict__ frob, uint8_t* __restrict__ ok) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ok[i] = g2_in_subgroup(pts[i], frob[0], frob[1]) ? 1 : 0;
}

__global__ void verify_unpack_kernel(VerifyKeys keys, const uint8_t* __restrict__ proofs, size_t stride, uint32_t n,
                                     G1Affine* __restrict__ P, G2Affine* __restrict__ Q, G1Affine* __restrict__ P2,
                                     G2Affine* __restrict__ Q2, G1Affine* __restrict__ commit, const Fp2* __restrict__ frob,
                                     uint32_t* __restrict__ bad) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* pr = proofs + (size_t)i * stride;
    const uint32_t role = blockIdx.y;
    uint32_t e = 0;
    if (role == 0) {
        G1Affine a;
        e = decompress_g1_point(pr, a);
        a.y = a.y.neg();
        P[4 * (size_t)i] = a;
        P[4 * (size_t)i + 1] = keys.alpha;
        Q[4 * (size_t)i + 1] = keys.beta2;
        Q[4 * (size_t)i + 2] = keys.gamma2;
        Q[4 * (size_t)i + 3] = keys.delta2;
        uint32_t cnt = ((uint32_t)pr[128] << 24) | ((uint32_t)pr[129] << 16) | ((uint32_t)pr[130] << 8) | pr[131];
        if (cnt != keys.n_commit) e |= 4;
        if (keys.n_commit) { Q2[2 * (size_t)i] = keys.ped_g; Q2[2 * (size_t)i + 1] = keys.ped_gneg; }
    } else if (role == 1) {
        G1Affine c;
        e = decompress_g1_point(pr + 96, c);
        P[4 * (size_t)i + 3] = c;
    } else if (role == 2) {
        G2Affine b;
        e = decompress_g2_point(pr + 32, b);
        if (!e && !g2_in_subgroup(b, frob[0], frob[1])) e = 8;   // on the twist but outside the r-torsion: gnark's decoder rejects it
        Q[4 * (size_t)i] = b;
    } else if (role == 3) {
        G1Affine c;
        e = decompress_g1_point(pr + 132, c);
        P2[2 * (size_t)i] = c;
        commit[i] = c;
    } else {
        G1Affine c;
        e = decompress_g1_point(pr + 164, c);
        P2[2 * (size_t)i + 1] = c;
    }
    if (e) atomicOr(&bad[i], e);
}
// kSum_i = sum_k pub_k K_k (the MSM result, XYZZ) + the commitment point, affine, into P[4i + 2]
__global__ void verify_ksum_kernel(const G1XYZZ* __restrict__ msm, const G1Affine* __restrict__ commit, uint32_t n,
                                   G1Affine* __restrict__ P) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1XYZZ v = msm[i];
    if (commit) v.madd(commit[i], false);
    P[4 * (size_t)i + 2] = v.to_affine();
}
// verdict = both pairing products are one and the proof was well-formed
__global__ void verify_verdict_kernel(const uint8_t* __restrict__ ok1, const uint8_t* __restrict__ ok2, const uint32_t* __restrict__ bad,
                                      uint32_t n, uint8_t* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = (ok1[i] && (!ok2 || ok2[i]) && !bad[i]) ? 1 : 0;
}

}  // namespace g16
