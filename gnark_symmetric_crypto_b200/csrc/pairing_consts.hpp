// Constants of the optimal ate pairing on BN254 (pure functions of the field characteristic p and the group order r;
// generated with Python big integers: (p - 1) / 3, (p - 1) / 6, the base-p digits of (p^4 - p^2 + 1) / r, and the loop count
// 6x + 2 for x = 4965661367192848881).
// Used by pairing.cuh. gnark-crypto v0.14.0 ecc/bn254/pairing.go computes the same map with a tower-field final
// exponentiation; the value compared against one is identical.
#pragma once
#include <cstdint>

namespace g16 {

__device__ const uint32_t PAIRING_PM1_DIV3[8] = {0x4829a9c2u, 0x69602eb2u, 0xcd7b4384u, 0xdd2b2385u, 0x808072c9u, 0xe81ac1e7u, 0xa065e00du, 0x10216f7bu};   // (p - 1) / 3
__device__ const uint32_t PAIRING_PM1_DIV6[8] = {0x2414d4e1u, 0x34b01759u, 0xe6bda1c2u, 0xee9591c2u, 0xc0403964u, 0xf40d60f3u, 0xd032f006u, 0x0810b7bdu};   // (p - 1) / 6
// hard part of the final exponentiation, h = (p^4 - p^2 + 1) / r, in base p: h = H[0] + H[1] p + H[2] p^2 + p^3 (digits < p, the
// top digit is 1), so that g^h = g^H0 * pi(g)^H1 * pi^2(g)^H2 * pi^3(g) is one 254-bit simultaneous exponentiation and one
// more product (pi = p-power Frobenius)
__device__ const uint32_t PAIRING_HARD_DIGITS[3][8] = {
    {0xd0f9fa91u, 0x85989436u, 0xfd736beau, 0x5cea24f6u, 0x3fd84104u, 0x048b6e19u, 0xe131a029u, 0x30644e72u},
    {0x606a30c5u, 0x138f3176u, 0xdae41fe4u, 0x3b852988u, 0x3fd84105u, 0x048b6e19u, 0xe131a029u, 0x30644e72u},
    {0xe87cfd46u, 0xf83e9682u, 0xeeb859fbu, 0x6f4d8248u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u},
};
static const uint64_t PAIRING_ATE_LOOP_LO = 0x9d797039be763ba8ULL;   // 6x + 2 = 2^64 + this
static const int PAIRING_STEPS = 64 + 36 + 2;            // doublings + additions of the loop + the two Frobenius additions

}  // namespace g16
