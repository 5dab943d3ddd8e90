// GLV constants of BN254 G1 (generated with Python big integers and checked against the oracle: phi(x, y) = (beta x, y) equals
// lambda * (x, y) on the generator). lambda^2 + lambda + 1 = 0 mod r, beta^3 = 1 mod p. Lattice basis (a1, b1), (a2, b2) with
// a_i + b_i lambda = 0 mod r, b1 < 0; g1 = floor(2^256 b2 / r), g2 = floor(2^256 |b1| / r). For k < r:
//   c1 = (k g1) >> 256, c2 = (k g2) >> 256, k1 = k - c1 a1 - c2 a2, k2 = c1 |b1| - c2 b2, k = k1 + k2 lambda mod r, |k1|, |k2| < 2^128.
// Same split as gnark-crypto v0.14.0 ecc.SplitScalar / ecc/bn254 G1Jac.mulGLV (used by G1Jac.ScalarMultiplication, which
// gnark's prove.go:174-295 calls for s*Ar and r*Bs1).
#pragma once
#include <cstdint>

namespace g16 {

__device__ const uint32_t GLV_BETA[8] = {0x77fffffeu, 0x57634731u, 0xacdb5c4fu, 0xd4f263f1u, 0xa0d48bacu, 0x59e26bceu, 0x00000000u, 0x00000000u};      // canonical
__device__ const uint32_t GLV_G1[3] = {0xc7e0b3d7u, 0xd91d232eu, 0x00000002u};
__device__ const uint32_t GLV_G2[5] = {0x391eb18du, 0x7a7bd9d4u, 0xa773d2cfu, 0x4ccef014u, 0x00000002u};
__device__ const uint32_t GLV_A1[2] = {0x94d213e3u, 0x89d32568u};
__device__ const uint32_t GLV_A2[4] = {0x1221250bu, 0x0be4e154u, 0xeeb859fdu, 0x6f4d8248u};
__device__ const uint32_t GLV_NB1[4] = {0x7d4f1128u, 0x8211bbebu, 0xeeb859fcu, 0x6f4d8248u};   // |b1|
__device__ const uint32_t GLV_B2[2] = {0x94d213e3u, 0x89d32568u};
static const int GLV_NIBBLES = 33;   // |k1|, |k2| < 2^128 in 200 000 random trials; worst case of the floor roundings < 2^129

}  // namespace g16
