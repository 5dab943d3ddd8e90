// Groth16 Setup on the GPU: trapdoor evaluation of the QAP on the host (a few hundred thousand field
// products), every group element of the proving and verifying key as a fixed-base product on the device, keys serialized in
// gnark's WriteTo layouts (SURVEY.md Appendices A and B).
//
// Replaces: keygen.go:359-435 (generateAES128 / generateAES256: groth16.Setup(r1cs) + pk.WriteTo / vk.WriteTo) and gnark
// v0.11.0 backend/groth16/bn254/setup.go (Setup: Lagrange evaluation at tau, A/B/C per wire, K split between pk and vk,
// BSB22 commitment keys) — the reference ships r1cs.aes128/256 and vk.aes128/256 but no pk.aes128/256.
// The trapdoor (tau, alpha, beta, gamma, delta and the Pedersen sigma) is an INPUT here, so a test can fix it and compare the
// key bytes with the oracle's Setup restatement; NULL draws it from the OS CSPRNG, which is what gnark does.
// Cold TU (-DG16_COLD).
#include "common.cuh"
#include "host_parse.hpp"
#include "prover_api.hpp"
#include "serialize.cuh"
#include <algorithm>
#include <set>

namespace g16 {

// out[i] = compressed(k_i * G), k canonical. Thread per scalar; the point is the G1 / G2 generator.
template <class C>
__global__ void setup_fixed_base_kernel(const Fr* __restrict__ k, uint32_t n, typename C::A gen, uint8_t* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    typename C::X r = scalar_mul(C::X::from_affine(gen), k[i]);
    typename C::A a = r.to_affine();
    if (sizeof(typename C::A) == sizeof(G1Affine)) g1_compress(*reinterpret_cast<const G1Affine*>(&a), out + (size_t)i * 32);
    else g2_compress(*reinterpret_cast<const G2Affine*>(&a), out + (size_t)i * 64);
}

static Fr fr_from_u64(uint64_t v) {
    Fr r = Fr::zero();
    r.l[0] = (uint32_t)v;
    r.l[1] = (uint32_t)(v >> 32);
    return r.to_mont();
}
static Fr fr_from_be(const uint8_t* b) {   // canonical big-endian (must be < r) -> Montgomery
    Fr v;
    for (int i = 0; i < 8; i++) {
        const uint8_t* q = b + 28 - 4 * i;
        v.l[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
    uint32_t t[8], m[8];
    for (int i = 0; i < 8; i++) m[i] = FrParams::mod(i);
    if (!sub8(t, v.l, m)) throw std::invalid_argument("setup: a trapdoor value is not reduced modulo the group order");
    return v.to_mont();
}
static void fr_to_be(const Fr& m, uint8_t* b) {
    Fr c = m.from_mont();
    for (int i = 0; i < 8; i++) {
        uint8_t* q = b + 28 - 4 * i;
        q[0] = (uint8_t)(c.l[i] >> 24); q[1] = (uint8_t)(c.l[i] >> 16); q[2] = (uint8_t)(c.l[i] >> 8); q[3] = (uint8_t)c.l[i];
    }
}
static Fr fr_pow_u64(Fr b, uint64_t e) {
    Fr r = Fr::one();
    while (e) {
        if (e & 1) r = r * b;
        b = b * b;
        e >>= 1;
    }
    return r;
}
static void put_be32(std::vector<uint8_t>& o, uint32_t v) { for (int s = 24; s >= 0; s -= 8) o.push_back((uint8_t)(v >> s)); }
static void put_be64(std::vector<uint8_t>& o, uint64_t v) { for (int s = 56; s >= 0; s -= 8) o.push_back((uint8_t)(v >> s)); }

// the 2^28-th root of unity gnark-crypto's fr/fft domain starts from (fr.Generator's 2-adic root)
static Fr root_2_28() {
    static const uint8_t be[32] = {0x2a, 0x3c, 0x09, 0xf0, 0xa5, 0x8a, 0x7e, 0x85, 0x00, 0xe0, 0xa7, 0xeb, 0x8e, 0xf6, 0x2a, 0xbc,
                                   0x40, 0x2d, 0x11, 0x1e, 0x41, 0x11, 0x2e, 0xd4, 0x9b, 0xd6, 0x1b, 0x6e, 0x72, 0x5b, 0x19, 0xf0};
    return fr_from_be(be);   // 19103219067921713944291392827692070036145651957329286315305642004821462161904
}

static Fp fp_small(uint32_t v) {
    Fp r = Fp::zero();
    r.l[0] = v;
    return r.to_mont();
}
static G1Affine g1_generator() { return {fp_small(1), fp_small(2)}; }   // gnark-crypto bn254 g1Gen = (1, 2)
// gnark-crypto bn254 g2Gen (the EIP-197 generator): X = x.a0 + x.a1 u, Y = y.a0 + y.a1 u, big-endian canonical
static G2Affine g2_generator() {
    static const uint8_t be[4][32] = {
        {0x18, 0x00, 0xde, 0xef, 0x12, 0x1f, 0x1e, 0x76, 0x42, 0x6a, 0x00, 0x66, 0x5e, 0x5c, 0x44, 0x79, 0x67, 0x43, 0x22, 0xd4, 0xf7, 0x5e, 0xda, 0xdd, 0x46, 0xde, 0xbd, 0x5c, 0xd9, 0x92, 0xf6, 0xed},
        {0x19, 0x8e, 0x93, 0x93, 0x92, 0x0d, 0x48, 0x3a, 0x72, 0x60, 0xbf, 0xb7, 0x31, 0xfb, 0x5d, 0x25, 0xf1, 0xaa, 0x49, 0x33, 0x35, 0xa9, 0xe7, 0x12, 0x97, 0xe4, 0x85, 0xb7, 0xae, 0xf3, 0x12, 0xc2},
        {0x12, 0xc8, 0x5e, 0xa5, 0xdb, 0x8c, 0x6d, 0xeb, 0x4a, 0xab, 0x71, 0x80, 0x8d, 0xcb, 0x40, 0x8f, 0xe3, 0xd1, 0xe7, 0x69, 0x0c, 0x43, 0xd3, 0x7b, 0x4c, 0xe6, 0xcc, 0x01, 0x66, 0xfa, 0x7d, 0xaa},
        {0x09, 0x06, 0x89, 0xd0, 0x58, 0x5f, 0xf0, 0x75, 0xec, 0x9e, 0x99, 0xad, 0x69, 0x0c, 0x33, 0x95, 0xbc, 0x4b, 0x31, 0x33, 0x70, 0xb3, 0x8e, 0xf3, 0x55, 0xac, 0xda, 0xdc, 0xd1, 0x22, 0x97, 0x5b},
    };
    G2Affine g;
    g.x.a0 = fp_from_be32(be[0], false); g.x.a1 = fp_from_be32(be[1], false);
    g.y.a0 = fp_from_be32(be[2], false); g.y.a1 = fp_from_be32(be[3], false);
    return g;
}

// k canonical (from_mont of the values) -> compressed points, in order
template <class C>
static std::vector<uint8_t> fixed_base(const std::vector<Fr>& mont_scalars, const typename C::A& gen, cudaStream_t st) {
    const size_t n = mont_scalars.size();
    const size_t sz = sizeof(typename C::A) == sizeof(G1Affine) ? 32 : 64;
    std::vector<uint8_t> out(n * sz);
    if (!n) return out;
    std::vector<Fr> can(n);
    for (size_t i = 0; i < n; i++) can[i] = mont_scalars[i].from_mont();
    DevBuf<Fr> dk;
    DevBuf<uint8_t> dout(n * sz);
    dk.upload(can.data(), n, st);
    auto k = setup_fixed_base_kernel<C>;
    G16_LAUNCH(k, div_up(n, 64), 64, 0, st, false, (const Fr*)dk.p, (uint32_t)n, gen, dout.p);
    G16_CHECK_LAUNCH();
    dout.download(out.data(), n * sz, st);
    G16_CUDA(cudaStreamSynchronize(st));
    return out;
}

// trapdoor_be: tau | alpha | beta | gamma | delta | sigma, 32-byte big-endian canonical each (all non-zero)
void setup_run(const uint8_t* r1cs_bytes, size_t r1cs_len, const uint8_t* trapdoor_be, std::vector<uint8_t>& pk, std::vector<uint8_t>& vk,
               cudaStream_t st) {
    R1csFile cs = parse_r1cs(r1cs_bytes, r1cs_len);
    const Fr tau = fr_from_be(trapdoor_be), alpha = fr_from_be(trapdoor_be + 32), beta = fr_from_be(trapdoor_be + 64),
             gamma = fr_from_be(trapdoor_be + 96), delta = fr_from_be(trapdoor_be + 128), sigma = fr_from_be(trapdoor_be + 160);
    if (tau.is_zero() || alpha.is_zero() || beta.is_zero() || gamma.is_zero() || delta.is_zero() || sigma.is_zero())
        throw std::invalid_argument("setup: trapdoor values must be non-zero");
    uint64_t n = 1;
    int lg = 0;
    while (n < cs.n_constraints) { n <<= 1; lg++; }
    if (lg > 28) throw std::invalid_argument("setup: constraint system too large for the 2-adicity of the field");
    const uint64_t nw = cs.n_wires();
    const Fr w = fr_pow_u64(root_2_28(), 1ull << (28 - lg));

    // ---- Lagrange basis at tau: L_j(tau) = w^j (tau^n - 1) / (n (tau - w^j)), one batched inversion
    std::vector<Fr> lag(n), wj(n), den(n), pref(n + 1);
    {
        const Fr tn1 = fr_pow_u64(tau, n) - Fr::one();
        if (tn1.is_zero()) throw std::invalid_argument("setup: tau is a root of the vanishing polynomial");
        const Fr zn = tn1 * fr_from_u64(n).inv();
        wj[0] = Fr::one();
        for (uint64_t j = 1; j < n; j++) wj[j] = wj[j - 1] * w;
        pref[0] = Fr::one();
        for (uint64_t j = 0; j < n; j++) { den[j] = tau - wj[j]; pref[j + 1] = pref[j] * den[j]; }
        Fr inv = pref[n].inv();
        for (uint64_t j = n; j-- > 0;) {
            lag[j] = wj[j] * zn * (inv * pref[j]);
            inv = inv * den[j];
        }
    }
    // ---- A_i(tau), B_i(tau), C_i(tau) per wire: every R1C term adds coeff * L_constraint(tau) to its wire (gnark setup.go)
    std::vector<Fr> A(nw, Fr::zero()), B(nw, Fr::zero()), Cc(nw, Fr::zero());
    {
        const std::vector<uint32_t>& cd = cs.calldata;
        const Fr* coeffs = reinterpret_cast<const Fr*>(cs.coeffs.data());
        const size_t ncoef = cs.coeffs.size() / 4;
        for (size_t i = 0; i < cs.n_instr(); i++) {
            if (cs.bp_kind[cs.bp_id[i]] != INS_R1C) continue;
            const uint64_t s0 = cs.start[i];
            if (s0 + 4 > cd.size() || s0 + cd[s0] > cd.size()) throw ParseError("r1cs: instruction calldata out of range");
            const uint32_t cnt[3] = {cd[s0 + 1], cd[s0 + 2], cd[s0 + 3]};
            if (cs.cons_off[i] >= cs.n_constraints) throw ParseError("r1cs: constraint offset out of range");
            const Fr lj = lag[cs.cons_off[i]];
            size_t pos = s0 + 4;
            std::vector<Fr>* side[3] = {&A, &B, &Cc};
            for (int sd = 0; sd < 3; sd++)
                for (uint32_t t = 0; t < cnt[sd]; t++) {
                    const uint32_t cid = cd[pos], wid = cd[pos + 1];
                    pos += 2;
                    if (cid >= ncoef) throw ParseError("r1cs: coefficient id out of range");
                    const uint32_t wire = wid == WIRE_CONST ? 0u : wid;   // constants sit on the ONE wire
                    if (wire >= nw) throw ParseError("r1cs: wire id out of range");
                    (*side[sd])[wire] = (*side[sd])[wire] + coeffs[cid] * lj;
                }
        }
    }
    const Fr dinv = delta.inv(), ginv = gamma.inv();
    std::set<uint32_t> committed, commit_wires;
    for (auto& ci : cs.commitments) {
        for (uint32_t x : ci.private_committed) committed.insert(x);
        commit_wires.insert((uint32_t)ci.commitment_index);
    }
    auto kval = [&](uint64_t i) { return beta * A[i] + alpha * B[i] + Cc[i]; };
    std::vector<Fr> sA, sB, sZ, sK, sVk;
    std::vector<uint8_t> inf_a(nw), inf_b(nw);
    uint64_t n_inf_a = 0, n_inf_b = 0;
    for (uint64_t i = 0; i < nw; i++) {
        inf_a[i] = A[i].is_zero(); inf_b[i] = B[i].is_zero();
        n_inf_a += inf_a[i]; n_inf_b += inf_b[i];
        if (!inf_a[i]) sA.push_back(A[i]);
        if (!inf_b[i]) sB.push_back(B[i]);
    }
    for (uint64_t i = cs.n_public; i < nw; i++)
        if (!committed.count((uint32_t)i) && !commit_wires.count((uint32_t)i)) sK.push_back(kval(i) * dinv);
    for (uint64_t i = 0; i < cs.n_public; i++) sVk.push_back(kval(i) * ginv);
    for (uint32_t i : commit_wires) sVk.push_back(kval(i) * ginv);   // std::set iterates in ascending wire order
    {
        // Z_i = tau^(brev i) (tau^n - 1) / delta, i < n - 1: gnark stores the quotient basis in the bit-reversed order its FFT leaves H in
        const Fr zt = (fr_pow_u64(tau, n) - Fr::one()) * dinv;
        std::vector<Fr> tp(n);
        tp[0] = Fr::one();
        for (uint64_t j = 1; j < n; j++) tp[j] = tp[j - 1] * tau;
        sZ.resize(n - 1);
        for (uint64_t i = 0; i + 1 < n; i++) {
            uint64_t r = 0;
            for (int b = 0; b < lg; b++) r |= ((i >> b) & 1ull) << (lg - 1 - b);
            sZ[i] = tp[r] * zt;
        }
    }
    const G1Affine g1 = g1_generator();
    const G2Affine g2 = g2_generator();
    std::vector<uint8_t> cA = fixed_base<G1>(sA, g1, st), cB = fixed_base<G1>(sB, g1, st), cZ = fixed_base<G1>(sZ, g1, st),
                         cK = fixed_base<G1>(sK, g1, st), cVk = fixed_base<G1>(sVk, g1, st), cB2 = fixed_base<G2>(sB, g2, st),
                         c_abd = fixed_base<G1>({alpha, beta, delta}, g1, st), c_bd2 = fixed_base<G2>({beta, delta}, g2, st),
                         c_g2 = fixed_base<G2>({gamma}, g2, st);
    auto append = [](std::vector<uint8_t>& o, const std::vector<uint8_t>& v) { o.insert(o.end(), v.begin(), v.end()); };
    auto append_n = [&](std::vector<uint8_t>& o, const std::vector<uint8_t>& v, size_t elem) { put_be32(o, (uint32_t)(v.size() / elem)); append(o, v); };

    // ---- proving key, SURVEY.md Appendix A
    pk.clear();
    put_be64(pk, n);
    {
        uint8_t b[32];
        const Fr five = fr_from_u64(5);
        const Fr hdr[5] = {fr_from_u64(n).inv(), w, w.inv(), five, five.inv()};
        for (const Fr& h : hdr) { fr_to_be(h, b); pk.insert(pk.end(), b, b + 32); }
    }
    pk.push_back(1);   // the domain's withPrecompute flag
    append(pk, c_abd);
    append_n(pk, cA, 32); append_n(pk, cB, 32); append_n(pk, cZ, 32); append_n(pk, cK, 32);
    append(pk, c_bd2);
    append_n(pk, cB2, 64);
    put_be64(pk, nw); put_be64(pk, n_inf_a); put_be64(pk, n_inf_b);
    append(pk, inf_a); append(pk, inf_b);
    put_be32(pk, (uint32_t)cs.commitments.size());
    std::vector<uint8_t> ped_vk;
    for (auto& ci : cs.commitments) {   // Pedersen proving key: Basis, BasisExpSigma; verifying key: G, G^(-1/sigma) on G2
        std::vector<Fr> basis, basis_sigma;
        for (uint32_t i : ci.private_committed) {
            if (i >= nw) throw ParseError("r1cs: committed wire out of range");
            basis.push_back(kval(i) * ginv);
            basis_sigma.push_back(basis.back() * sigma);
        }
        append_n(pk, fixed_base<G1>(basis, g1, st), 32);
        append_n(pk, fixed_base<G1>(basis_sigma, g1, st), 32);
        append(ped_vk, fixed_base<G2>({Fr::one(), sigma.inv().neg()}, g2, st));
    }
    // ---- verifying key, SURVEY.md Appendix B: alpha1 beta1 beta2 gamma2 delta1 delta2 | K | commitment info | Pedersen vks
    vk.clear();
    vk.insert(vk.end(), c_abd.begin(), c_abd.begin() + 64);
    vk.insert(vk.end(), c_bd2.begin(), c_bd2.begin() + 64);
    append(vk, c_g2);
    vk.insert(vk.end(), c_abd.begin() + 64, c_abd.end());
    vk.insert(vk.end(), c_bd2.begin() + 64, c_bd2.end());
    append_n(vk, cVk, 32);
    put_be32(vk, (uint32_t)cs.commitments.size());
    for (auto& ci : cs.commitments) {
        put_be32(vk, (uint32_t)ci.public_and_commitment_committed.size());
        for (uint32_t x : ci.public_and_commitment_committed) put_be64(vk, x);
    }
    put_be32(vk, (uint32_t)cs.commitments.size());
    append(vk, ped_vk);
}

}  // namespace g16
