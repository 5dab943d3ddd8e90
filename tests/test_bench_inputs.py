"""CPU: bench.py's self-contained input generator (hashlib + numpy ChaCha20, no oracle import in the GPU arm's process) is the
same stream as the tests' oracle-based generator, and both bench arms print one `config`."""
import hashlib

import bench
from conftest import batch_inputs


def test_bench_generator_matches_the_tests_generator(oracle):
    for n, seed in ((1, b"g16-b200-batch"), (37, b"g16-b200-batch"), (8, b"g16-b200-batch-rank3")):
        assert bench.make_requests(n, seed) == batch_inputs(n, seed)
    # RFC 7539 block function against the oracle's, block counters 0..4
    key = hashlib.sha256(b"k").digest()
    assert bench._chacha20_blocks(key, 5) == b"".join(oracle.chacha20_block(key, i, bytes(12)) for i in range(5))
    assert bench.R_MOD == oracle.R_MOD


def test_both_arms_print_the_same_config():
    for world in (1, 2, 8):
        assert bench.config_dict(world) == bench.config_dict(world)
        assert set(bench.config_dict(world)) == {"workload", "batch_per_gpu", "inputs", "l2", "parallelism"}
    src = open(bench.__file__).read()
    assert src.count('"config": config_dict(') == 2      # the GPU arm and the reference arm


def test_gpu_arm_does_not_import_the_oracle_for_inputs():
    src = open(bench.__file__).read()
    gpu_arm = src[src.index("def run_gpu(args):"):src.index("def main():")]
    # the only oracle use left in the GPU arm's process is the cpu_baseline leg (cpu_reference_run), which is the checker's job
    assert "from conftest" not in src and "import conftest" not in src
    assert "from oracle" not in gpu_arm and "import oracle" not in gpu_arm
