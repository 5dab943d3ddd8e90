"""CPU: the MSM paths that only switch on at production sizes — the batch-affine pairwise levels (csrc/msm_ba.cuh, from 2^21
entries) and the per-row shared-memory counting sort (csrc/msm.cuh msm_rowsort_kernel, from 32 rows) — forced onto the small
emulation cases. The switches are read once per process, so each variant runs tests/test_emu.py -k msm in a child process."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("env", [
    {"G16_MSM_BA_MIN": "1", "G16_MSM_BA_K": "1", "G16_MSM_TREE_BLOCK": "0"},
    {"G16_MSM_BA_MIN": "1", "G16_MSM_BA_K": "3", "G16_BA_INV2_MIN": "1"},
    {"G16_MSM_BA_MIN": "1", "G16_MSM_BA_K": "2", "G16_MSM_BA_LEFT": "3"},
    {"G16_MSM_ROWSORT": "2", "G16_MSM_BA_MIN": "1", "G16_MSM_BA_K": "2"},
], ids=["batch_affine_k1_serial_tree", "batch_affine_k3_two_level_inversion", "batch_affine_k2_leftovers_3", "rowsort_batch_affine_k2"])
def test_msm_variants_on_emulation(emu, env):
    out = subprocess.run([sys.executable, "-m", "pytest", "tests/test_emu.py", "-q", "-x", "-k", "msm", "-p", "no:cacheprovider"],
                         capture_output=True, text=True, env=dict(os.environ, **env), cwd=str(ROOT), timeout=1200)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert " passed" in out.stdout
