// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is linked into, imported by or executed from the
// product (gnark_symmetric_crypto_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs use it, and only as the checker / CPU baseline.
//
// CPU restatement of the BN254 arithmetic that the reference reaches through
//   github.com/consensys/gnark-crypto v0.14.0 (go.mod:9; source NOT on this box):
//   ecc/bn254/fp, ecc/bn254/fr (4x64-bit-limb Montgomery, R = 2^256), internal/fptower E2.
// Call sites in the reference: libraries/prover/impl/provers.go:148,216 (groth16.Prove).
// Algorithms are the published ones (CIOS Montgomery multiplication, Fermat inversion,
// p = 3 mod 4 square roots); constants are SURVEY.md Appendix G and are recomputed at start-up, not trusted.
#pragma once
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <cstdlib>

typedef uint64_t u64;
typedef unsigned __int128 u128;

struct FieldParams {
    u64 M[4];    // modulus
    u64 INV;     // -M^-1 mod 2^64
    u64 R1[4];   // 2^256 mod M   (Montgomery one)
    u64 R2[4];   // 2^512 mod M
    u64 PM2[4];  // M - 2
    void init(const u64 m[4]);
};

static inline int cmp4(const u64 a[4], const u64 b[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] < b[i]) return -1;
        if (a[i] > b[i]) return 1;
    }
    return 0;
}
static inline u64 add4(u64 r[4], const u64 a[4], const u64 b[4]) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a[i] + b[i]; r[i] = (u64)c; c >>= 64; }
    return (u64)c;
}
static inline u64 sub4(u64 r[4], const u64 a[4], const u64 b[4]) {
    u64 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a[i] - b[i] - br;
        r[i] = (u64)t;
        br = (u64)(t >> 64) & 1;
    }
    return br;
}

inline void FieldParams::init(const u64 m[4]) {
    memcpy(M, m, 32);
    // Newton iteration for the inverse of m[0] mod 2^64
    u64 x = 1;
    for (int i = 0; i < 7; i++) x *= 2 - m[0] * x;
    INV = (u64)0 - x;
    // R1 = 2^256 mod M by 256 modular doublings of 1 ; R2 by 256 more
    u64 v[4] = {1, 0, 0, 0};
    for (int i = 0; i < 512; i++) {
        u64 t[4];
        u64 c = add4(t, v, v);
        if (c || cmp4(t, M) >= 0) sub4(t, t, M);
        memcpy(v, t, 32);
        if (i == 255) memcpy(R1, v, 32);
    }
    memcpy(R2, v, 32);
    u64 two[4] = {2, 0, 0, 0};
    sub4(PM2, M, two);
}

extern FieldParams g_fp, g_fr;

// Generic 4-limb Montgomery element; Tag selects the modulus.
struct FpTag { static inline const FieldParams& P() { return g_fp; } };
struct FrTag { static inline const FieldParams& P() { return g_fr; } };

template <class Tag>
struct Fe {
    u64 l[4];
    static inline const FieldParams& P() { return Tag::P(); }
    static Fe zero() { Fe r; memset(r.l, 0, 32); return r; }
    static Fe one() { Fe r; memcpy(r.l, P().R1, 32); return r; }
    bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
    bool operator==(const Fe& o) const { return memcmp(l, o.l, 32) == 0; }
    bool operator!=(const Fe& o) const { return !(*this == o); }
    Fe operator+(const Fe& o) const {
        Fe r;
        u64 c = add4(r.l, l, o.l);
        if (c || cmp4(r.l, P().M) >= 0) sub4(r.l, r.l, P().M);
        return r;
    }
    Fe operator-(const Fe& o) const {
        Fe r;
        if (sub4(r.l, l, o.l)) add4(r.l, r.l, P().M);
        return r;
    }
    Fe neg() const {
        if (is_zero()) return *this;
        Fe r;
        sub4(r.l, P().M, l);
        return r;
    }
    Fe dbl() const { return *this + *this; }
    // CIOS Montgomery product: a*b*2^-256 mod M
    Fe operator*(const Fe& o) const {
        const FieldParams& p = P();
        u64 t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; i++) {
            u128 c = 0;
            for (int j = 0; j < 4; j++) {
                c += (u128)l[j] * o.l[i] + t[j];
                t[j] = (u64)c;
                c >>= 64;
            }
            c += t[4];
            t[4] = (u64)c;
            t[5] = (u64)(c >> 64);
            u64 m = t[0] * p.INV;
            c = ((u128)m * p.M[0] + t[0]) >> 64;
            for (int j = 1; j < 4; j++) {
                c += (u128)m * p.M[j] + t[j];
                t[j - 1] = (u64)c;
                c >>= 64;
            }
            c += t[4];
            t[3] = (u64)c;
            t[4] = t[5] + (u64)(c >> 64);
        }
        Fe r;
        memcpy(r.l, t, 32);
        if (t[4] || cmp4(r.l, p.M) >= 0) sub4(r.l, r.l, p.M);
        return r;
    }
    Fe sqr() const { return *this * *this; }
    Fe pow(const u64 e[4]) const {
        Fe r = one();
        for (int i = 255; i >= 0; i--) {
            r = r.sqr();
            if ((e[i / 64] >> (i % 64)) & 1) r = r * *this;
        }
        return r;
    }
    Fe pow_u64(u64 e) const { u64 ee[4] = {e, 0, 0, 0}; return pow(ee); }
    Fe inv() const { return pow(P().PM2); }   // 0 -> 0, like gnark's Inverse
    // canonical (non-Montgomery) little-endian limbs
    void to_canon(u64 out[4]) const {
        Fe o; memset(o.l, 0, 32); o.l[0] = 1;
        Fe r = *this * o;
        memcpy(out, r.l, 32);
    }
    static Fe from_canon(const u64 in[4]) {
        Fe a, r2;
        memcpy(a.l, in, 32);
        memcpy(r2.l, P().R2, 32);
        return a * r2;
    }
    static Fe from_u64(u64 v) { u64 c[4] = {v, 0, 0, 0}; return from_canon(c); }
    // big-endian 32-byte canonical encodings (gnark file/wire format)
    void to_be(uint8_t out[32]) const {
        u64 c[4]; to_canon(c);
        for (int i = 0; i < 4; i++) for (int b = 0; b < 8; b++) out[31 - (i * 8 + b)] = (uint8_t)(c[i] >> (8 * b));
    }
    static Fe from_be(const uint8_t in[32]) {
        u64 c[4] = {0, 0, 0, 0};
        for (int i = 0; i < 4; i++) for (int b = 0; b < 8; b++) c[i] |= (u64)in[31 - (i * 8 + b)] << (8 * b);
        return from_canon(c);
    }
    // "lexicographically largest": canonical value > (M-1)/2
    bool lex_largest() const {
        u64 c[4]; to_canon(c);
        u64 h[4];   // (M-1)/2
        const u64* m = P().M;
        for (int i = 0; i < 4; i++) h[i] = (m[i] >> 1) | (i < 3 ? (m[i + 1] << 63) : 0);
        return cmp4(c, h) > 0;
    }
};

typedef Fe<FpTag> Fp;
typedef Fe<FrTag> Fr;

// Fp2 = Fp[u]/(u^2+1)   (gnark-crypto internal/fptower E2{A0,A1})
struct Fp2 {
    Fp a0, a1;
    static Fp2 zero() { return {Fp::zero(), Fp::zero()}; }
    static Fp2 one() { return {Fp::one(), Fp::zero()}; }
    bool is_zero() const { return a0.is_zero() && a1.is_zero(); }
    bool operator==(const Fp2& o) const { return a0 == o.a0 && a1 == o.a1; }
    bool operator!=(const Fp2& o) const { return !(*this == o); }
    Fp2 operator+(const Fp2& o) const { return {a0 + o.a0, a1 + o.a1}; }
    Fp2 operator-(const Fp2& o) const { return {a0 - o.a0, a1 - o.a1}; }
    Fp2 neg() const { return {a0.neg(), a1.neg()}; }
    Fp2 dbl() const { return {a0.dbl(), a1.dbl()}; }
    Fp2 conj() const { return {a0, a1.neg()}; }
    Fp2 operator*(const Fp2& o) const {
        Fp t0 = a0 * o.a0, t1 = a1 * o.a1;
        Fp t2 = (a0 + a1) * (o.a0 + o.a1);
        return {t0 - t1, t2 - t0 - t1};
    }
    Fp2 mul_fp(const Fp& s) const { return {a0 * s, a1 * s}; }
    Fp2 sqr() const { return *this * *this; }
    Fp2 inv() const {
        Fp n = (a0.sqr() + a1.sqr()).inv();
        return {a0 * n, (a1 * n).neg()};
    }
    Fp2 pow(const u64* e, int nlimbs) const {
        Fp2 r = one();
        for (int i = nlimbs * 64 - 1; i >= 0; i--) {
            r = r.sqr();
            if ((e[i / 64] >> (i % 64)) & 1) r = r * *this;
        }
        return r;
    }
    // gnark-crypto E2 ordering for compression: compare A1 first, then A0
    bool lex_largest() const {
        if (a1.is_zero()) return a0.lex_largest();
        return a1.lex_largest();
    }
};
