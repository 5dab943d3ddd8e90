#!/bin/bash
# Sanitizer evidence for the kernels' index arithmetic and block synchronisation. compute-sanitizer is
# refused on the GPU pool ("closed on this pool and stays closed"), so the same kernel sources run here, compiled as C++ by
# tests/emu (-DG16_EMU), under AddressSanitizer + UBSan (out-of-bounds / misaligned / overflow in scatter, merge, tile code)
# and ThreadSanitizer (kernels that use __syncthreads run as real host threads: shared-memory races).
#   bash scripts/emu_sanitize.sh asan|tsan [pytest -k expression] [tag]
# e.g. G16_MSM_BA_MIN=1 G16_MSM_BA_K=3 bash scripts/emu_sanitize.sh asan msm ba   (the batch-affine levels forced onto the small cases)
set -e
cd "$(dirname "$0")/.."
SAN=${1:-asan}
KEXPR=${2:-"msm or ntt or field or group or pairing or aes_witness or solver or compute_h or subgroup or verifier"}
SYSCXX=/usr/bin/g++
make -C tests/emu SAN=$SAN CXX=$SYSCXX -j"$(nproc)" -s
RT=$($SYSCXX -print-file-name=lib${SAN}.so)
mkdir -p profiles
export G16_EMU_SO=tests/emu/_build_${SAN}/libg16emu.so
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
export TSAN_OPTIONS="halt_on_error=1 report_signal_unsafe=0 exitcode=66"   # a race report ends the run with a failure
TAG=${3:+_$3}
OUT=profiles/emu_${SAN}${TAG}_r02.txt
{
  echo "# tests/test_emu.py -k '$KEXPR' against $G16_EMU_SO ($SAN build of the product's kernel sources, $(date -u +%FT%TZ))"
  echo "# env: G16_MSM_BA_MIN=${G16_MSM_BA_MIN:-} G16_MSM_BA_K=${G16_MSM_BA_K:-}"
  LD_PRELOAD="$RT $( $SYSCXX -print-file-name=libstdc++.so.6)" python -m pytest tests/test_emu.py -x -q -s -k "$KEXPR" -p no:cacheprovider 2>&1 | grep -v "^\[msm\]" | tail -40; echo "sanitizer reports: $(grep -c -E "ThreadSanitizer|AddressSanitizer|runtime error" "$OUT" || true)"
} > "$OUT" 2>&1 || true
tail -15 "$OUT"
