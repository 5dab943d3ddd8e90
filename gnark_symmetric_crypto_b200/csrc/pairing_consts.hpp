// Constants of the optimal ate pairing on BN254 (pure functions of the field characteristic p and the group order r;
// generated with Python big integers: e = (p^12 - 1) / r, (p - 1) / 3, and the loop count 6x + 2 for x = 4965661367192848881).
// Used by pairing.cuh. gnark-crypto v0.14.0 ecc/bn254/pairing.go computes the same map with a tower-field final
// exponentiation; the value compared against one is identical.
#pragma once
#include <cstdint>

namespace g16 {

static const int PAIRING_FINAL_EXP_LIMBS = 88;   // little-endian 32-bit limbs, 2790 bits
__device__ const uint32_t PAIRING_FINAL_EXP[PAIRING_FINAL_EXP_LIMBS] = {
    0xca86f120u, 0x86964b64u, 0xe54523a4u, 0x40a4efb7u, 0x96e84abbu, 0x837fa978u, 0xb9b2b918u, 0x361102b6u,
    0xf35692dau, 0xc0de81deu, 0xa6c3c760u, 0xbe04c7e8u, 0xd570bb7fu, 0xd766f9c9u, 0x83561841u, 0xc230974du,
    0xc3be69a3u, 0x5bba1668u, 0x10526294u, 0x7f3811c4u, 0xdadda71cu, 0x29baee7du, 0x145da900u, 0xbf813b8du,
    0x423f9a2cu, 0x641bbadfu, 0x44eacc5eu, 0xa80bb4eau, 0x14fde37cu, 0xcd656648u, 0x580291d2u, 0x4a0364b9u,
    0x0826f0ddu, 0xee93dfb1u, 0xc5514724u, 0x6b42db8du, 0x0b0f3785u, 0xbb10cf43u, 0x6f804216u, 0x40494e40u,
    0xacf3aafbu, 0x55cfe107u, 0xe0ebae87u, 0x2088ec80u, 0x11a337a0u, 0x846a3ed0u, 0x1e3a5195u, 0x48a45a4au,
    0xdfc50e16u, 0xe5664568u, 0x4c0cc4ebu, 0xab6a4129u, 0xd268c7dau, 0x82d0d602u, 0xed3cc48au, 0x6668449au,
    0xb2015dfcu, 0x5062cd0fu, 0xb1ddb3d1u, 0x7f2940a8u, 0x2a226448u, 0x77f5b63au, 0x61e443aeu, 0xfef07813u,
    0x88d5c6c8u, 0xf977870eu, 0x1f676baau, 0x790364a6u, 0xceaddea3u, 0x5887e72eu, 0xa09a1b70u, 0x1377e563u,
    0x1bd8c3b2u, 0x0c54efeeu, 0xd524d8f7u, 0x3ec3d15au, 0xb2383a5du, 0xdaf15466u, 0xbb94fec0u, 0xe1e30a73u,
    0x5f3f7be2u, 0x6a1c7101u, 0x6369b1ffu, 0x842d43bfu, 0x107d20bcu, 0x20fddadfu, 0x4b6dc970u, 0x0000002fu,
};
__device__ const uint32_t PAIRING_PM1_DIV3[8] = {0x4829a9c2u, 0x69602eb2u, 0xcd7b4384u, 0xdd2b2385u, 0x808072c9u, 0xe81ac1e7u, 0xa065e00du, 0x10216f7bu};   // (p - 1) / 3
static const uint64_t PAIRING_ATE_LOOP_LO = 0x9d797039be763ba8ULL;   // 6x + 2 = 2^64 + this
static const int PAIRING_STEPS = 64 + 36 + 2;            // doublings + additions of the loop + the two Frobenius additions

}  // namespace g16
