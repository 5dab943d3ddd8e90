// Big-endian field-element and compressed-point (de)serialisation shared by the key-loading kernels (prover_kernels.cuh)
// and the proof assembly (assemble.cuh). Replaces gnark-crypto v0.14.0 ecc/bn254/marshal.go (G1Affine/G2Affine Bytes,
// the 0x80 / 0xC0 / 0x40 flag bits) as used by gnark Proof.WriteTo (marshal.go:32-59) and ProvingKey.ReadFrom (:311-348).
#pragma once
#include "common.cuh"

namespace g16 {

// *noncanonical (optional) is set when the encoded integer is >= p: gnark-crypto's point decoders read coordinates with
// fp.Element.SetBytesCanonical and reject such encodings (x and x + p would otherwise decode to the same point, i.e. two
// byte strings for one proof); to_mont alone would reduce them silently.
FD Fp fp_from_be32(const uint8_t* b, bool mask_flags, bool* noncanonical = nullptr) {
    Fp v;
    for (int i = 0; i < 8; i++) {
        const uint8_t* q = b + 28 - 4 * i;
        v.l[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    }
    if (mask_flags) v.l[7] &= 0x3FFFFFFFu;
    if (noncanonical) {
        uint32_t t[8], m[8];
        for (int i = 0; i < 8; i++) m[i] = FpParams::mod(i);
        if (!sub8(t, v.l, m)) *noncanonical = true;   // no borrow <=> v >= p
    }
    return v.to_mont();
}
FD void fp_to_be32(const Fp& m, uint8_t* b) {
    Fp c = m.from_mont();
    for (int i = 0; i < 8; i++) {
        uint8_t* q = b + 28 - 4 * i;
        q[0] = (uint8_t)(c.l[i] >> 24); q[1] = (uint8_t)(c.l[i] >> 16); q[2] = (uint8_t)(c.l[i] >> 8); q[3] = (uint8_t)c.l[i];
    }
}
FD void g1_compress(const G1Affine& p, uint8_t* out) {
    if (p.is_inf()) { for (int i = 0; i < 32; i++) out[i] = 0; out[0] = 0x40; return; }
    fp_to_be32(p.x, out);
    out[0] |= p.y.lex_largest() ? 0xC0 : 0x80;
}
FD void g2_compress(const G2Affine& p, uint8_t* out) {
    if (p.is_inf()) { for (int i = 0; i < 64; i++) out[i] = 0; out[0] = 0x40; return; }
    fp_to_be32(p.x.a1, out);
    fp_to_be32(p.x.a0, out + 32);
    out[0] |= p.y.lex_largest() ? 0xC0 : 0x80;
}

}  // namespace g16
