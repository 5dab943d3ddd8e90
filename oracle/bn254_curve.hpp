// ORACLE — TEST INFRASTRUCTURE ONLY (see bn254_field.hpp header).
//
// CPU restatement of gnark-crypto v0.14.0 ecc/bn254 G1/G2 group law, point (de)compression
// (reference reads these through prove_impl.go:86-91 ProvingKey.ReadFrom) and
// multi-exponentiation ((*G1Jac).MultiExp / (*G2Jac).MultiExp, called by groth16 Prove — provers.go:148,216).
// Formulas: EFD "dbl-2009-l", "add-2007-bl" (a = 0). Results are compared as AFFINE points, which are unique,
// so the choice of projective formulas does not affect parity.
#pragma once
#include "bn254_field.hpp"
#include <vector>
#include <thread>

template <class F>
struct Aff {
    F x, y;   // (0,0) = point at infinity (gnark convention)
    bool is_inf() const { return x.is_zero() && y.is_zero(); }
    Aff neg() const { return {x, y.neg()}; }
};

template <class F>
struct Jac {
    F X, Y, Z;
    static Jac inf() { return {F::one(), F::one(), F::zero()}; }
    bool is_inf() const { return Z.is_zero(); }
    static Jac from_aff(const Aff<F>& a) {
        if (a.is_inf()) return inf();
        return {a.x, a.y, F::one()};
    }
    Jac neg() const { return {X, Y.neg(), Z}; }
    Jac dbl() const {
        if (is_inf()) return *this;
        F A = X.sqr(), B = Y.sqr(), C = B.sqr();
        F D = ((X + B).sqr() - A - C).dbl();
        F E = A.dbl() + A;
        F Fq = E.sqr();
        Jac r;
        r.X = Fq - D.dbl();
        r.Y = E * (D - r.X) - C.dbl().dbl().dbl();
        r.Z = (Y * Z).dbl();
        return r;
    }
    Jac add(const Jac& o) const {
        if (is_inf()) return o;
        if (o.is_inf()) return *this;
        F Z1Z1 = Z.sqr(), Z2Z2 = o.Z.sqr();
        F U1 = X * Z2Z2, U2 = o.X * Z1Z1;
        F S1 = Y * o.Z * Z2Z2, S2 = o.Y * Z * Z1Z1;
        if (U1 == U2) {
            if (S1 == S2) return dbl();
            return inf();
        }
        F H = U2 - U1;
        F I = H.dbl().sqr();
        F J = H * I;
        F r = (S2 - S1).dbl();
        F V = U1 * I;
        Jac out;
        out.X = r.sqr() - J - V.dbl();
        out.Y = r * (V - out.X) - (S1 * J).dbl();
        out.Z = ((Z + o.Z).sqr() - Z1Z1 - Z2Z2) * H;
        return out;
    }
    Jac add_aff(const Aff<F>& a) const { return add(from_aff(a)); }
    Aff<F> to_aff() const {
        if (is_inf()) return {F::zero(), F::zero()};
        F zi = Z.inv();
        F zi2 = zi.sqr();
        return {X * zi2, Y * zi2 * zi};
    }
    // scalar given as canonical little-endian limbs
    Jac mul(const u64* k, int nlimbs) const {
        Jac r = inf();
        for (int i = nlimbs * 64 - 1; i >= 0; i--) {
            r = r.dbl();
            if ((k[i / 64] >> (i % 64)) & 1) r = r.add(*this);
        }
        return r;
    }
};

typedef Aff<Fp> G1A;
typedef Jac<Fp> G1J;
typedef Aff<Fp2> G2A;
typedef Jac<Fp2> G2J;

extern Fp g_b1;    // 3
extern Fp2 g_b2;   // 3/(9+u)

static inline bool g1_on_curve(const G1A& p) {
    if (p.is_inf()) return true;
    return p.y.sqr() == p.x.sqr() * p.x + g_b1;
}
static inline bool g2_on_curve(const G2A& p) {
    if (p.is_inf()) return true;
    return p.y.sqr() == p.x.sqr() * p.x + g_b2;
}

bool fp_sqrt(const Fp& a, Fp& out);
bool fp2_sqrt(const Fp2& a, Fp2& out);

// gnark-crypto compressed encodings
int g1_decompress(const uint8_t in[32], G1A& out);
int g2_decompress(const uint8_t in[64], G2A& out);
void g1_compress(const G1A& p, uint8_t out[32]);
void g2_compress(const G2A& p, uint8_t out[64]);

// Pippenger bucket method with signed digits; scalars canonical LE limbs (4 x u64 each).
template <class F>
Jac<F> msm_pippenger(const Aff<F>* pts, const u64* scalars, size_t n, int nthreads);
template <class F>
Jac<F> msm_naive(const Aff<F>* pts, const u64* scalars, size_t n);
