// BN254 Fp / Fr arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form with R = 2^256, i.e. the SAME bytes as
// gnark-crypto's in-memory fp.Element / fr.Element (4 x u64 little-endian limbs), so buffers cross the C-ABI unchanged.
//
// Replaces: gnark-crypto v0.14.0 ecc/bn254/fp/element_ops_amd64.s (fp.mul, ADX mulx/adcx/adox),
// fp/element.go:436-447 (Sub), fr/element_ops_amd64.s (fr.mul) — reached from libraries/prover/impl/provers.go:148,216.
//
// Multiplication is an interleaved (CIOS) Montgomery product written as mad.lo.cc / madc.hi.cc carry chains.
// ptxas fuses every lo/hi pair on the same operands into one IMAD.WIDE.U32(.X) with the carry in a predicate
// (checked with cuobjdump -sass), so a product costs 64 wide MADs + 64 for the reduction + 8 IMAD for the quotients.
// To keep each chain free of carry hazards the running value is split into two accumulators whose 64-bit product slots
// are aligned to even resp. odd limb positions (E: limbs 0..7, O: limbs 1..8); after each round the value is shifted
// one limb and the two accumulators swap roles.
//
// G16_EMU: the same algorithms compiled as plain C++ (tests/emu) — every asm block has a C twin with identical semantics.
#pragma once
#include <cstdint>

#if defined(G16_EMU)
#include "cuemu.h"
#define G16_ASM 0
#else
#include <cuda_runtime.h>
#if defined(__CUDA_ARCH__)
#define G16_ASM 1
#else
#define G16_ASM 0
#endif
#endif

#define FD __host__ __device__ __forceinline__
// Cold translation units (-DG16_COLD: key loading, proof assembly, G2) keep the 180-instruction Montgomery product out
// of line so that ptxas finishes in seconds; the hot kernels (G1 bucket accumulation, NTT) inline it.
// The G2 MSM unit (-DG16_COLD_FP2) sits in between: the Fp product is inlined, but the Fp2 product and square that contain it
// (3 and 2 Fp products) stay out of line — one call per 3 products instead of one per product, and ptxas still only sees a
// G2 addition as ten calls (the fully inlined Fp2 group law did not finish compiling in 25 minutes).
#if defined(G16_COLD_FP2) && !defined(G16_EMU)
#define FD_MUL FD
#define FD_MUL2 __host__ __device__ __noinline__
#elif defined(G16_COLD) && !defined(G16_EMU)
#define FD_MUL __host__ __device__ __noinline__
#define FD_MUL2 FD
#else
#define FD_MUL FD
#define FD_MUL2 FD
#endif

namespace g16 {

struct FpParams {
    static FD constexpr uint32_t mod(int i) {
        switch (i) {
            case 0: return 0xd87cfd47u; case 1: return 0x3c208c16u; case 2: return 0x6871ca8du; case 3: return 0x97816a91u;
            case 4: return 0x8181585du; case 5: return 0xb85045b6u; case 6: return 0xe131a029u; default: return 0x30644e72u;
        }
    }
    static FD constexpr uint32_t inv() { return 0xe4866389u; }   // -p^-1 mod 2^32
    // R mod p (Montgomery one) and R^2 mod p, SURVEY.md Appendix G
    static FD constexpr uint32_t one(int i) {
        switch (i) {
            case 0: return 0xc58f0d9du; case 1: return 0xd35d438du; case 2: return 0xf5c70b3du; case 3: return 0x0a78eb28u;
            case 4: return 0x7879462cu; case 5: return 0x666ea36fu; case 6: return 0x9a07df2fu; default: return 0x0e0a77c1u;
        }
    }
    static FD constexpr uint32_t r2(int i) {
        switch (i) {
            case 0: return 0x538afa89u; case 1: return 0xf32cfc5bu; case 2: return 0xd44501fbu; case 3: return 0xb5e71911u;
            case 4: return 0x0a417ff6u; case 5: return 0x47ab1effu; case 6: return 0xcab8351fu; default: return 0x06d89f71u;
        }
    }
};
struct FrParams {
    static FD constexpr uint32_t mod(int i) {
        switch (i) {
            case 0: return 0xf0000001u; case 1: return 0x43e1f593u; case 2: return 0x79b97091u; case 3: return 0x2833e848u;
            case 4: return 0x8181585du; case 5: return 0xb85045b6u; case 6: return 0xe131a029u; default: return 0x30644e72u;
        }
    }
    static FD constexpr uint32_t inv() { return 0xefffffffu; }
    static FD constexpr uint32_t one(int i) {
        switch (i) {
            case 0: return 0x4ffffffbu; case 1: return 0xac96341cu; case 2: return 0x9f60cd29u; case 3: return 0x36fc7695u;
            case 4: return 0x7879462eu; case 5: return 0x666ea36fu; case 6: return 0x9a07df2fu; default: return 0x0e0a77c1u;
        }
    }
    static FD constexpr uint32_t r2(int i) {
        switch (i) {
            case 0: return 0xae216da7u; case 1: return 0x1bb8e645u; case 2: return 0xe35c59e3u; case 3: return 0x53fe3ab1u;
            case 4: return 0x53bb8085u; case 5: return 0x8c49833du; case 6: return 0x7f4e44a5u; default: return 0x0216d0b1u;
        }
    }
};

// ---------------------------------------------------------------------------------------------------------
// carry-chain building blocks
// ---------------------------------------------------------------------------------------------------------

// acc[2k+1]:acc[2k] = x_k * b   (four independent 64-bit products)
FD void mul4(uint32_t acc[8], uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b) {
#if G16_ASM
    asm("mul.lo.u32 %0, %8, %12;\n\tmul.hi.u32 %1, %8, %12;\n\t"
        "mul.lo.u32 %2, %9, %12;\n\tmul.hi.u32 %3, %9, %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\tmul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\tmul.hi.u32 %7, %11, %12;"
        : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]), "=r"(acc[7])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
#else
    const uint32_t x[4] = {x0, x1, x2, x3};
    for (int k = 0; k < 4; k++) {
        uint64_t p = (uint64_t)x[k] * b;
        acc[2 * k] = (uint32_t)p;
        acc[2 * k + 1] = (uint32_t)(p >> 32);
    }
#endif
}

// acc (8 limbs, four 64-bit slots) += (x0,x1,x2,x3) * b, one carry chain; returns the carry out of the top limb
FD uint32_t mad4(uint32_t acc[8], uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b) {
    uint32_t cout;
#if G16_ASM
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]),
          "=r"(cout)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
#else
    const uint32_t x[4] = {x0, x1, x2, x3};
    uint64_t carry = 0;
    for (int k = 0; k < 4; k++) {
        unsigned __int128 t = (unsigned __int128)((uint64_t)x[k] * b) + (((uint64_t)acc[2 * k + 1] << 32) | acc[2 * k]) + carry;
        acc[2 * k] = (uint32_t)t;
        acc[2 * k + 1] = (uint32_t)(t >> 32);
        carry = (uint64_t)(t >> 64);
    }
    cout = (uint32_t)carry;
#endif
    return cout;
}

// The one-limb right shift that makes the accumulators swap roles:
//   lo0 += e[1]                                   (carry feeds the chain below)
//   t    = (x0,x1,x2,x3) * b + (e >> 64) + carry  (e[2..7] land in t[0..5])
FD void mad4_shift(uint32_t t[8], uint32_t& lo0, const uint32_t e[8], uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3,
                   uint32_t b) {
#if G16_ASM
    asm("add.cc.u32 %8, %8, %9;\n\t"
        "madc.lo.cc.u32 %0, %16, %20, %10;\n\tmadc.hi.cc.u32 %1, %16, %20, %11;\n\t"
        "madc.lo.cc.u32 %2, %17, %20, %12;\n\tmadc.hi.cc.u32 %3, %17, %20, %13;\n\t"
        "madc.lo.cc.u32 %4, %18, %20, %14;\n\tmadc.hi.cc.u32 %5, %18, %20, %15;\n\t"
        "madc.lo.cc.u32 %6, %19, %20, 0;\n\tmadc.hi.u32 %7, %19, %20, 0;"
        : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "+r"(lo0)
        : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(x0), "r"(x1), "r"(x2), "r"(x3),
          "r"(b));
#else
    const uint32_t x[4] = {x0, x1, x2, x3};
    uint64_t s0 = (uint64_t)lo0 + e[1];
    lo0 = (uint32_t)s0;
    uint64_t carry = s0 >> 32;
    for (int k = 0; k < 4; k++) {
        uint64_t add = 0;
        if (k < 3) add = ((uint64_t)e[2 * k + 3] << 32) | e[2 * k + 2];
        unsigned __int128 v = (unsigned __int128)((uint64_t)x[k] * b) + add + carry;
        t[2 * k] = (uint32_t)v;
        t[2 * k + 1] = (uint32_t)(v >> 32);
        carry = (uint64_t)(v >> 64);
    }
#endif
}

// ---- variants for the dedicated squaring (Fe::sqr): the multiplicand of round i has no limbs below i, so the first K of the
// four 64-bit slots of a block hold no product. For the chain without carry-in (mad4) the skipped slots vanish; for the
// shifted chain (mad4_shift) they only pass the carry along (two ALU additions instead of two multiplier operations).
template <int K>
FD uint32_t mad4_from(uint32_t acc[8], uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b) {
    static_assert(K >= 0 && K <= 4, "K");
    if (K == 0) return mad4(acc, x0, x1, x2, x3, b);
    if (K == 4) return 0u;
#if G16_ASM
    uint32_t cout;
    if (K == 1) {
        asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\tmadc.hi.cc.u32 %1, %7, %10, %1;\n\t"
            "madc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\t"
            "madc.lo.cc.u32 %4, %9, %10, %4;\n\tmadc.hi.cc.u32 %5, %9, %10, %5;\n\t"
            "addc.u32 %6, 0, 0;"
            : "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "=r"(cout)
            : "r"(x1), "r"(x2), "r"(x3), "r"(b));
    } else if (K == 2) {
        asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\tmadc.hi.cc.u32 %1, %5, %7, %1;\n\t"
            "madc.lo.cc.u32 %2, %6, %7, %2;\n\tmadc.hi.cc.u32 %3, %6, %7, %3;\n\t"
            "addc.u32 %4, 0, 0;"
            : "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "=r"(cout)
            : "r"(x2), "r"(x3), "r"(b));
    } else {
        asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\t"
            "addc.u32 %2, 0, 0;"
            : "+r"(acc[6]), "+r"(acc[7]), "=r"(cout)
            : "r"(x3), "r"(b));
    }
    return cout;
#else
    return mad4(acc, K > 0 ? 0u : x0, K > 1 ? 0u : x1, K > 2 ? 0u : x2, K > 3 ? 0u : x3, b);
#endif
}
template <int K>
FD void mad4_shift_from(uint32_t t[8], uint32_t& lo0, const uint32_t e[8], uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3,
                        uint32_t b) {
    static_assert(K >= 0 && K <= 3, "K");
    if (K == 0) { mad4_shift(t, lo0, e, x0, x1, x2, x3, b); return; }
#if G16_ASM
    if (K == 1) {
        asm("add.cc.u32 %8, %8, %9;\n\t"
            "addc.cc.u32 %0, %10, 0;\n\taddc.cc.u32 %1, %11, 0;\n\t"
            "madc.lo.cc.u32 %2, %16, %19, %12;\n\tmadc.hi.cc.u32 %3, %16, %19, %13;\n\t"
            "madc.lo.cc.u32 %4, %17, %19, %14;\n\tmadc.hi.cc.u32 %5, %17, %19, %15;\n\t"
            "madc.lo.cc.u32 %6, %18, %19, 0;\n\tmadc.hi.u32 %7, %18, %19, 0;"
            : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "+r"(lo0)
            : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(x1), "r"(x2), "r"(x3), "r"(b));
    } else if (K == 2) {
        asm("add.cc.u32 %8, %8, %9;\n\t"
            "addc.cc.u32 %0, %10, 0;\n\taddc.cc.u32 %1, %11, 0;\n\t"
            "addc.cc.u32 %2, %12, 0;\n\taddc.cc.u32 %3, %13, 0;\n\t"
            "madc.lo.cc.u32 %4, %16, %18, %14;\n\tmadc.hi.cc.u32 %5, %16, %18, %15;\n\t"
            "madc.lo.cc.u32 %6, %17, %18, 0;\n\tmadc.hi.u32 %7, %17, %18, 0;"
            : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "+r"(lo0)
            : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(x2), "r"(x3), "r"(b));
    } else {
        asm("add.cc.u32 %8, %8, %9;\n\t"
            "addc.cc.u32 %0, %10, 0;\n\taddc.cc.u32 %1, %11, 0;\n\t"
            "addc.cc.u32 %2, %12, 0;\n\taddc.cc.u32 %3, %13, 0;\n\t"
            "addc.cc.u32 %4, %14, 0;\n\taddc.cc.u32 %5, %15, 0;\n\t"
            "madc.lo.cc.u32 %6, %16, %17, 0;\n\tmadc.hi.u32 %7, %16, %17, 0;"
            : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "+r"(lo0)
            : "r"(e[1]), "r"(e[2]), "r"(e[3]), "r"(e[4]), "r"(e[5]), "r"(e[6]), "r"(e[7]), "r"(x3), "r"(b));
    }
#else
    mad4_shift(t, lo0, e, K > 0 ? 0u : x0, K > 1 ? 0u : x1, K > 2 ? 0u : x2, x3, b);
#endif
}

// r = a + b (8 limbs), returns carry
FD uint32_t add8(uint32_t r[8], const uint32_t a[8], const uint32_t b[8]) {
    uint32_t c;
#if G16_ASM
    asm("add.cc.u32 %0, %9, %17;\n\taddc.cc.u32 %1, %10, %18;\n\taddc.cc.u32 %2, %11, %19;\n\taddc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\taddc.cc.u32 %5, %14, %22;\n\taddc.cc.u32 %6, %15, %23;\n\taddc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint64_t cc = 0;
    for (int i = 0; i < 8; i++) { cc += (uint64_t)a[i] + b[i]; r[i] = (uint32_t)cc; cc >>= 32; }
    c = (uint32_t)cc;
#endif
    return c;
}
// r = a - b (8 limbs), returns borrow as 0 / 0xffffffff
FD uint32_t sub8(uint32_t r[8], const uint32_t a[8], const uint32_t b[8]) {
    uint32_t bw;
#if G16_ASM
    asm("sub.cc.u32 %0, %9, %17;\n\tsubc.cc.u32 %1, %10, %18;\n\tsubc.cc.u32 %2, %11, %19;\n\tsubc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\tsubc.cc.u32 %5, %14, %22;\n\tsubc.cc.u32 %6, %15, %23;\n\tsubc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(bw)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint64_t br = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t t = (uint64_t)a[i] - b[i] - br;
        r[i] = (uint32_t)t;
        br = (t >> 32) & 1;
    }
    bw = br ? 0xffffffffu : 0u;
#endif
    return bw;
}

// ---------------------------------------------------------------------------------------------------------
// field element
// ---------------------------------------------------------------------------------------------------------
template <class P>
struct alignas(16) Fe {
    uint32_t l[8];

    static FD Fe zero() { Fe r; for (int i = 0; i < 8; i++) r.l[i] = 0; return r; }
    static FD Fe one() { Fe r; for (int i = 0; i < 8; i++) r.l[i] = P::one(i); return r; }
    static FD Fe r2() { Fe r; for (int i = 0; i < 8; i++) r.l[i] = P::r2(i); return r; }
    static FD Fe modulus() { Fe r; for (int i = 0; i < 8; i++) r.l[i] = P::mod(i); return r; }
    FD bool is_zero() const { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= l[i]; return o == 0; }
    FD bool operator==(const Fe& b) const { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= l[i] ^ b.l[i]; return o == 0; }
    FD bool operator!=(const Fe& b) const { return !(*this == b); }

    // conditional subtraction of the modulus: v in [0, 2p) -> [0, p)
    FD void reduce_once() {
        uint32_t t[8], m[8];
        for (int i = 0; i < 8; i++) m[i] = P::mod(i);
        uint32_t bw = sub8(t, l, m);
        for (int i = 0; i < 8; i++) l[i] = bw ? l[i] : t[i];
    }
    friend FD Fe operator+(const Fe& a, const Fe& b) {
        Fe r;
        add8(r.l, a.l, b.l);   // a,b < p < 2^254: no carry out
        r.reduce_once();
        return r;
    }
    friend FD Fe operator-(const Fe& a, const Fe& b) {
        Fe r;
        uint32_t bw = sub8(r.l, a.l, b.l);
        uint32_t m[8];
        for (int i = 0; i < 8; i++) m[i] = P::mod(i) & bw;
        add8(r.l, r.l, m);
        return r;
    }
    FD Fe neg() const { return is_zero() ? *this : modulus_minus(*this); }
    static FD Fe modulus_minus(const Fe& a) { Fe r, m = modulus(); sub8(r.l, m.l, a.l); return r; }
    FD Fe dbl() const { return *this + *this; }

    // Montgomery product a*b*2^-256 mod p, inputs and output in [0,p)
    friend FD_MUL Fe operator*(const Fe& a, const Fe& b) {
        uint32_t E[8], O[8], t[8];
        // round 0
        mul4(E, a.l[0], a.l[2], a.l[4], a.l[6], b.l[0]);
        mul4(O, a.l[1], a.l[3], a.l[5], a.l[7], b.l[0]);
        uint32_t m = E[0] * P::inv();
        mad4(O, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
        O[7] += mad4(E, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
#pragma unroll
        for (int i = 1; i < 8; i++) {
            // value = E + O*2^32 with E[0] == 0. Shift one limb: O becomes the even accumulator, E>>64 seeds the odd one.
            mad4_shift(t, O[0], E, a.l[1], a.l[3], a.l[5], a.l[7], b.l[i]);
            t[7] += mad4(O, a.l[0], a.l[2], a.l[4], a.l[6], b.l[i]);
            m = O[0] * P::inv();
            mad4(t, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
            t[7] += mad4(O, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
#pragma unroll
            for (int k = 0; k < 8; k++) { E[k] = O[k]; O[k] = t[k]; }
        }
        // result = (E >> 32) + O
        Fe r;
        uint32_t sh[8];
        for (int k = 0; k < 7; k++) sh[k] = E[k + 1];
        sh[7] = 0;
        add8(r.l, O, sh);
        r.reduce_once();
        return r;
    }
    // Dedicated Montgomery squaring. a^2 = sum_i a_i 2^(32 i) * C_i with C_i = a_i 2^(32 i) + 2 (a div 2^(32 (i+1))) 2^(32 (i+1)):
    // round i of the interleaved product multiplies a_i by a vector that has no limbs below i — the symmetric partial products
    // are taken once, doubled — so 36 of the 64 operand products remain (the 64 of the reduction stay): 116 wide multiply-adds
    // instead of 144. The limbs of C_i: a_i at i, a_(i+1) << 1 at i+1 (its top bit lives in the next limb), the limbs of 2a above.
    // Same result as a * a (the value is (a^2 + M p) / 2^256 either way), bit for bit; G16_NO_SQR falls back to the product.
#if defined(G16_NO_SQR)
    FD Fe sqr() const { return *this * *this; }
#else
    template <int I>
    static FD void sqr_round(uint32_t E[8], uint32_t O[8], const uint32_t a[8], const uint32_t e[8], const uint32_t d[8]) {
        // multiplicand limb j of round I
#define G16_SQ_C(j) ((j) < I ? 0u : ((j) == I ? a[(j)] : ((j) == I + 1 ? e[(j)] : d[(j)])))
        uint32_t t[8];
        mad4_shift_from<I / 2>(t, O[0], E, G16_SQ_C(1), G16_SQ_C(3), G16_SQ_C(5), G16_SQ_C(7), a[I]);
        t[7] += mad4_from<(I + 1) / 2>(O, G16_SQ_C(0), G16_SQ_C(2), G16_SQ_C(4), G16_SQ_C(6), a[I]);
#undef G16_SQ_C
        uint32_t m = O[0] * P::inv();
        mad4(t, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
        t[7] += mad4(O, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
#pragma unroll
        for (int k = 0; k < 8; k++) { E[k] = O[k]; O[k] = t[k]; }
    }
    FD_MUL Fe sqr() const {
        uint32_t e[8], d[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            e[j] = l[j] << 1;
            d[j] = j ? (e[j] | (l[j - 1] >> 31)) : e[j];
        }
        uint32_t E[8], O[8];
        // round 0: C_0 = (a_0, a_1 << 1, limbs 2.. of 2a)
        mul4(E, l[0], d[2], d[4], d[6], l[0]);
        mul4(O, e[1], d[3], d[5], d[7], l[0]);
        uint32_t m = E[0] * P::inv();
        mad4(O, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
        O[7] += mad4(E, P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
        sqr_round<1>(E, O, l, e, d);
        sqr_round<2>(E, O, l, e, d);
        sqr_round<3>(E, O, l, e, d);
        sqr_round<4>(E, O, l, e, d);
        sqr_round<5>(E, O, l, e, d);
        sqr_round<6>(E, O, l, e, d);
        sqr_round<7>(E, O, l, e, d);
        Fe r;
        uint32_t sh[8];
        for (int k = 0; k < 7; k++) sh[k] = E[k + 1];
        sh[7] = 0;
        add8(r.l, O, sh);
        r.reduce_once();
        return r;
    }
#endif

    // Montgomery <-> canonical
    FD Fe to_mont() const { return *this * r2(); }
    FD Fe from_mont() const {
        Fe o = zero();
        o.l[0] = 1;
        return *this * o;
    }
    // this^e, e given as 8 canonical 32-bit limbs (uniform across the warp: no divergence)
    FD Fe pow(const uint32_t e[8]) const {
        Fe r = one();
        for (int i = 255; i >= 0; i--) {
            r = r.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1) r = r * *this;
        }
        return r;
    }
    FD Fe inv_fermat() const {   // this^(p-2); 0 -> 0
        uint32_t e[8];
        for (int i = 0; i < 8; i++) e[i] = P::mod(i);
        e[0] -= 2;   // low limbs of both moduli are > 2
        return pow(e);
    }
    // Inverse by the binary extended Euclidean algorithm (shifts, subtractions, no products): on this GPU a dependent
    // Montgomery product costs ~0.6 us, so the 380-product Fermat chain above is ~230 us of latency wherever an inversion
    // sits on a critical path (proof assembly, the affine Miller loop of the verifier, divisions of the AES solver); this
    // loop is a few thousand cheap instructions. 0 -> 0, like gnark's fp/fr.Element.Inverse.
    // The loop runs on the Montgomery representative aR and yields (aR)^-1; one product by R^3 turns that into a^-1 R.
    FD Fe inv() const {
        if (is_zero()) return *this;
        uint32_t u[8], v[8], x1[8], x2[8], m[8], t[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { u[i] = l[i]; v[i] = m[i] = P::mod(i); x1[i] = i == 0 ? 1u : 0u; x2[i] = 0u; }
#define G16_HALVE(a, top)                                                                   \
        do {                                                                                \
            _Pragma("unroll") for (int i_ = 0; i_ < 7; i_++) a[i_] = (a[i_] >> 1) | (a[i_ + 1] << 31); \
            a[7] = (a[7] >> 1) | ((top) << 31);                                             \
        } while (0)
#define G16_IS_ONE(a) (a[0] == 1u && (a[1] | a[2] | a[3] | a[4] | a[5] | a[6] | a[7]) == 0u)
        while (!G16_IS_ONE(u) && !G16_IS_ONE(v)) {
            while (!(u[0] & 1u)) {
                G16_HALVE(u, 0u);
                uint32_t c = 0;
                if (x1[0] & 1u) c = add8(x1, x1, m);
                G16_HALVE(x1, c);
            }
            while (!(v[0] & 1u)) {
                G16_HALVE(v, 0u);
                uint32_t c = 0;
                if (x2[0] & 1u) c = add8(x2, x2, m);
                G16_HALVE(x2, c);
            }
            if (!sub8(t, u, v)) {   // u >= v
#pragma unroll
                for (int i = 0; i < 8; i++) u[i] = t[i];
                if (sub8(x1, x1, x2)) add8(x1, x1, m);
            } else {
                sub8(v, v, u);
                if (sub8(x2, x2, x1)) add8(x2, x2, m);
            }
        }
        Fe y;
        const bool from_u = G16_IS_ONE(u);
#pragma unroll
        for (int i = 0; i < 8; i++) y.l[i] = from_u ? x1[i] : x2[i];
#undef G16_HALVE
#undef G16_IS_ONE
        return y * (r2() * r2());   // R^2 * R^2 * R^-1 = R^3 ; y * R^3 * R^-1 = (aR)^-1 R^2 = a^-1 R
    }
    // canonical value > (p-1)/2 ?  ("lexicographically largest")
    FD bool lex_largest() const {
        Fe c = from_mont();
        uint32_t h[8], t[8];
        for (int i = 0; i < 8; i++) h[i] = (P::mod(i) >> 1) | (i < 7 ? (P::mod(i + 1) << 31) : 0u);
        return sub8(t, h, c.l) != 0;   // h - c borrows  <=>  c > h
    }
};

typedef Fe<FpParams> Fp;
typedef Fe<FrParams> Fr;

// Fp2 = Fp[u]/(u^2+1)  (gnark-crypto internal/fptower E2; a17 "fptower.mulAdxE2")
struct Fp2 {
    Fp a0, a1;
    static FD Fp2 zero() { return {Fp::zero(), Fp::zero()}; }
    static FD Fp2 one() { return {Fp::one(), Fp::zero()}; }
    FD bool is_zero() const { return a0.is_zero() && a1.is_zero(); }
    FD bool operator==(const Fp2& b) const { return a0 == b.a0 && a1 == b.a1; }
    FD bool operator!=(const Fp2& b) const { return !(*this == b); }
    friend FD Fp2 operator+(const Fp2& a, const Fp2& b) { return {a.a0 + b.a0, a.a1 + b.a1}; }
    friend FD Fp2 operator-(const Fp2& a, const Fp2& b) { return {a.a0 - b.a0, a.a1 - b.a1}; }
    FD Fp2 neg() const { return {a0.neg(), a1.neg()}; }
    FD Fp2 dbl() const { return {a0.dbl(), a1.dbl()}; }
    FD Fp2 conj() const { return {a0, a1.neg()}; }
    friend FD_MUL2 Fp2 operator*(const Fp2& a, const Fp2& b) {   // Karatsuba, 3 Fp products
        Fp t0 = a.a0 * b.a0, t1 = a.a1 * b.a1;
        Fp t2 = (a.a0 + a.a1) * (b.a0 + b.a1);
        return {t0 - t1, t2 - t0 - t1};
    }
    FD_MUL2 Fp2 sqr() const {   // (a0+a1)(a0-a1), 2 a0 a1
        Fp s = a0 + a1, d = a0 - a1, p = a0 * a1;
        return {s * d, p.dbl()};
    }
    FD Fp2 mul_fp(const Fp& s) const { return {a0 * s, a1 * s}; }
    FD Fp2 inv() const {
        Fp n = (a0.sqr() + a1.sqr()).inv();
        return {a0 * n, (a1 * n).neg()};
    }
    FD Fp2 pow(const uint32_t e[8]) const {
        Fp2 r = one();
        for (int i = 255; i >= 0; i--) {
            r = r.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1) r = r * *this;
        }
        return r;
    }
    FD bool lex_largest() const { return a1.is_zero() ? a0.lex_largest() : a1.lex_largest(); }
};

}  // namespace g16
