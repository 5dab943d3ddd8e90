// BN254 G1 / G2 group law for the MSM kernels: affine inputs, extended-Jacobian "XYZZ" accumulators
// (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; infinity <=> ZZ == 0). Templated on the coordinate field (Fp for G1, Fp2 for G2).
//
// Replaces: gnark-crypto v0.14.0 ecc/bn254/g1.go:747-,833- (g1JacExtended.add / addMixed),
// g2.go:728-,814-, and the batch-affine bucket adds of multiexp_affine.go — reached from provers.go:148,216.
// Formulas: EFD madd-2008-s, add-2008-s, dbl-2008-s-1 (a = 0). Every exceptional case (infinity, P+P, P+(-P)) is
// handled exactly, because the final affine point must be bit-identical to gnark's.
#pragma once
#include "field.cuh"

namespace g16 {

template <class F>
struct Affine {
    F x, y;   // (0,0) = infinity (gnark convention)
    FD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    static FD Affine inf() { return {F::zero(), F::zero()}; }
};

template <class F>
struct XYZZ {
    F X, Y, ZZ, ZZZ;
    static FD XYZZ inf() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
    FD bool is_inf() const { return ZZ.is_zero(); }
    static FD XYZZ from_affine(const Affine<F>& p) {
        if (p.is_inf()) return inf();
        return {p.x, p.y, F::one(), F::one()};
    }
    FD XYZZ neg() const { return {X, Y.neg(), ZZ, ZZZ}; }

    // 2 * affine point (mdbl-2008-s-1)
    static FD XYZZ dbl_affine(const Affine<F>& p) {
        if (p.is_inf()) return inf();
        F U = p.y.dbl();
        F V = U.sqr();
        F W = U * V;
        F S = p.x * V;
        F xx = p.x.sqr();
        F M = xx.dbl() + xx;
        XYZZ r;
        r.X = M.sqr() - S.dbl();
        r.Y = M * (S - r.X) - W * p.y;
        r.ZZ = V;
        r.ZZZ = W;
        return r;
    }
    FD XYZZ dbl() const {   // dbl-2008-s-1
        if (is_inf()) return *this;
        F U = Y.dbl();
        F V = U.sqr();
        F W = U * V;
        F S = X * V;
        F xx = X.sqr();
        F M = xx.dbl() + xx;
        XYZZ r;
        r.X = M.sqr() - S.dbl();
        r.Y = M * (S - r.X) - W * Y;
        r.ZZ = V * ZZ;
        r.ZZZ = W * ZZZ;
        return r;
    }
    // this += p (affine), `negate` flips the sign of p (signed-digit buckets)
    FD void madd(const Affine<F>& p_in, bool negate) {
        if (p_in.is_inf()) return;
        Affine<F> p = p_in;
        if (negate) p.y = p.y.neg();
        if (is_inf()) {
            X = p.x; Y = p.y; ZZ = F::one(); ZZZ = F::one();
            return;
        }
        F U2 = p.x * ZZ;
        F S2 = p.y * ZZZ;
        F P = U2 - X;
        F R = S2 - Y;
        if (P.is_zero()) {
            if (R.is_zero()) *this = dbl_affine(p);
            else *this = inf();
            return;
        }
        F PP = P.sqr();
        F PPP = P * PP;
        F Q = X * PP;
        F X3 = R.sqr() - PPP - Q.dbl();
        Y = R * (Q - X3) - Y * PPP;
        X = X3;
        ZZ = ZZ * PP;
        ZZZ = ZZZ * PPP;
    }
    // this += o (add-2008-s)
    FD void add(const XYZZ& o) {
        if (o.is_inf()) return;
        if (is_inf()) { *this = o; return; }
        F U1 = X * o.ZZ;
        F U2 = o.X * ZZ;
        F S1 = Y * o.ZZZ;
        F S2 = o.Y * ZZZ;
        F P = U2 - U1;
        F R = S2 - S1;
        if (P.is_zero()) {
            if (R.is_zero()) *this = dbl();
            else *this = inf();
            return;
        }
        F PP = P.sqr();
        F PPP = P * PP;
        F Q = U1 * PP;
        F X3 = R.sqr() - PPP - Q.dbl();
        Y = R * (Q - X3) - S1 * PPP;
        X = X3;
        ZZ = ZZ * o.ZZ * PP;
        ZZZ = ZZZ * o.ZZZ * PPP;
    }
    // affine form (one inversion): x = X/ZZ, y = Y/ZZZ.  1/ZZ = ZZ^2 / ZZZ^2 * ... computed from a single inverse of ZZZ:
    // ZZ^3 = ZZZ^2  =>  1/ZZ = ZZ^2/ZZZ^2.
    FD Affine<F> to_affine() const {
        if (is_inf()) return Affine<F>::inf();
        F zi3 = ZZZ.inv();            // 1/ZZZ
        F zi2 = (ZZ * zi3).sqr();     // (ZZ/ZZZ)^2 = 1/ZZ   since ZZZ^2 = ZZ^3
        return {X * zi2, Y * zi3};
    }
};

// k * P, k = 8 little-endian 32-bit limbs (double-and-add, MSB first). The limbs are passed BY VALUE in a struct and
// indexed with compile-time constants only (outer loop unrolled): a dynamically indexed thread-local scalar array plus
// an operand passed by reference was miscompiled by nvcc 12.9 / ptxas for sm_100a in a cold translation unit
// (observed on B200: the scalar bits read back garbage), so neither appears here.
struct Scalar256 {
    uint32_t w[8];
};
template <class F>
FD XYZZ<F> scalar_mul(XYZZ<F> p, Scalar256 k) {
    XYZZ<F> r = XYZZ<F>::inf();
#pragma unroll
    for (int wi = 7; wi >= 0; wi--) {
        uint32_t word = k.w[wi];
#pragma unroll 1
        for (int b = 31; b >= 0; b--) {
            r = r.dbl();
            if ((word >> b) & 1u) r.add(p);
        }
    }
    return r;
}
template <class F, class P>
FD XYZZ<F> scalar_mul(const XYZZ<F>& p, const Fe<P>& k) {
    Scalar256 s;
    for (int i = 0; i < 8; i++) s.w[i] = k.l[i];
    return scalar_mul(p, s);
}

typedef Affine<Fp> G1Affine;
typedef Affine<Fp2> G2Affine;
typedef XYZZ<Fp> G1XYZZ;
typedef XYZZ<Fp2> G2XYZZ;

// group traits used as template arguments of the kernels
struct G1 {
    typedef Fp F;
    typedef Affine<Fp> A;
    typedef XYZZ<Fp> X;
};
struct G2 {
    typedef Fp2 F;
    typedef Affine<Fp2> A;
    typedef XYZZ<Fp2> X;
};

}  // namespace g16
