"""ctypes binding of libg16b200.so (include/g16b200.h).

The product path loads ONLY the nvcc-built CUDA library `gnark_symmetric_crypto_b200/lib/libg16b200.so` and raises if it
is missing — there is no CPU fallback. (`bind()` is also used by tests/emu to bind the test-only host-emulation build of
the same sources; that library is never reachable through this package's public API.)
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "lib" / "libg16b200.so"

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)


class GoSlice(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_longlong), ("cap", C.c_longlong)]


class ProveReturn(C.Structure):
    _fields_ = [("r0", C.c_void_p), ("r1", C.c_longlong)]


EXPORTS = {
    # name: (restype, argtypes)
    "g16_version": (C.c_int, []),
    "g16_last_error": (C.c_char_p, []),
    "g16_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "g16_init": (C.c_int, [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "g16_free": (None, [C.c_void_p]),
    "g16_init_multi": (C.c_int, [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.c_size_t, C.POINTER(C.c_void_p)]),
    "g16_ctx_devices": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.c_size_t, C.POINTER(C.c_size_t)]),
    "g16_ctx_device_handle": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "g16_last_batch_status": (C.c_int, [C.c_void_p, u32p, C.c_size_t]),
    "g16_info": (C.c_int, [C.c_void_p, u64p]),
    "g16_prove_witness": (C.c_int, [C.c_void_p, u64p, C.c_size_t, u8p, u8p, C.POINTER(C.c_size_t)]),
    "g16_prove_chacha_batch": (C.c_int, [C.c_void_p, C.c_size_t, u8p, u8p, u32p, u8p, u8p, u8p, u8p]),
    "g16_chacha_batch_stage": (C.c_int, [C.c_void_p, C.c_size_t, u8p, u8p, u32p, u8p, u8p]),
    "g16_chacha_batch_run": (C.c_int, [C.c_void_p, f32p]),
    "g16_chacha_batch_fetch": (C.c_int, [C.c_void_p, u8p, u8p]),
    "g16_prove_aes_batch": (C.c_int, [C.c_void_p, C.c_size_t, u8p, C.c_size_t, u8p, u32p, u8p, u8p, u8p, u8p]),
    "g16_aes_batch_stage": (C.c_int, [C.c_void_p, C.c_size_t, u8p, C.c_size_t, u8p, u32p, u8p, u8p]),
    "g16_solve_ex": (C.c_int, [C.c_void_p, u64p, C.c_size_t, C.c_size_t, u8p, u64p, u64p, u64p, u64p]),
    "g16_aes_witness": (C.c_int, [u8p, C.c_size_t, u8p, u32p, u8p, C.c_size_t, u8p, u64p]),
    "g16_bsb22_challenge": (C.c_int, [u64p, C.c_size_t, u64p]),
    "g16_setup": (C.c_int, [C.c_char_p, C.c_size_t, u8p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "g16_verify_init": (C.c_int, [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "g16_verify_info": (C.c_int, [C.c_void_p, u64p]),
    "g16_verify_batch": (C.c_int, [C.c_void_p, C.c_size_t, u8p, C.c_void_p, C.c_int, u8p, C.POINTER(C.c_float)]),
    "g16_verify_free": (None, [C.c_void_p]),
    "g16_g2_subgroup_check": (C.c_int, [u64p, C.c_size_t, u8p]),
    "g16_pairing_check": (C.c_int, [u64p, u64p, C.c_size_t, C.c_size_t, u8p]),
    "g16_set_schedule": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "g16_last_stage_ms": (C.c_int, [C.c_void_p, f32p]),
    "g16_last_counters": (C.c_int, [C.c_void_p, u64p]),
    "g16_last_counters_ex": (C.c_int, [C.c_void_p, u64p]),
    "g16_field_op": (C.c_int, [C.c_int, C.c_int, u64p, u64p, u64p, C.c_size_t]),
    "g16_group_op": (C.c_int, [C.c_int, C.c_int, u64p, u64p, u64p, C.c_size_t]),
    "g16_decompress": (C.c_int, [C.c_int, u8p, u64p, C.c_size_t]),
    "g16_msm": (C.c_int, [C.c_int, u64p, u64p, C.c_int, C.c_size_t, C.c_int, u64p, f32p]),
    "g16_bitq_sum": (C.c_int, [C.c_int, u64p, C.c_size_t, C.c_size_t, u64p, C.c_size_t, u64p, C.POINTER(C.c_uint32)]),
    "g16_msm_plan_create": (C.c_int, [C.c_int, u64p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "g16_msm_plan_set_scalars": (C.c_int, [C.c_void_p, u64p, C.c_int]),
    "g16_msm_plan_precompute": (C.c_int, [C.c_void_p, C.c_int]),
    "g16_msm_plan_run": (C.c_int, [C.c_void_p, u64p, f32p]),
    "g16_msm_plan_free": (None, [C.c_void_p]),
    "g16_ntt": (C.c_int, [u64p, C.c_size_t, C.c_int, C.c_int, f32p]),
    "g16_ntt_bench": (C.c_int, [C.c_size_t, C.c_size_t, C.c_int, f32p, u64p]),
    "g16_compute_h": (C.c_int, [C.c_void_p, u64p, u64p, u64p, u64p]),
    "g16_solve": (C.c_int, [C.c_void_p, u64p, C.c_size_t, C.c_size_t, u64p, u64p, u64p, u64p]),
    "g16_prove_witness_detail": (C.c_int, [C.c_void_p, u64p, C.c_size_t, u8p, u8p, C.POINTER(C.c_size_t), u64p, u64p, u64p]),
    "g16_imad_peak": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "g16_imad_chain_rate": (C.c_int, [C.POINTER(C.c_double)]),
    "enforce_binding": (None, []),
    "InitAlgorithm": (C.c_ubyte, [C.c_ubyte, GoSlice, GoSlice]),
    "Free": (None, [C.c_void_p]),
    "Prove": (ProveReturn, [GoSlice]),
    "ProveBatch": (ProveReturn, [GoSlice]),
    "g16_libprove_stats": (C.c_int, [C.c_int, u64p]),
    "InitVerifier": (C.c_ubyte, [C.c_ubyte, GoSlice]),
    "Verify": (C.c_ubyte, [GoSlice]),
}


def bind(path) -> C.CDLL:
    lib = C.CDLL(str(path))
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return lib


_LIB = None


def load() -> C.CDLL:
    """The CUDA library, or an ImportError — never a fallback."""
    global _LIB
    if _LIB is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). gnark_symmetric_crypto_b200 has no CPU fallback.")
        _LIB = bind(LIB_PATH)
    return _LIB
