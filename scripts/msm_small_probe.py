"""Latency probe of the standalone G1 MSM at small sizes (BASELINE config 5, 2^16 .. 2^20): one-shot and fixed-base plans,
stage times from the library's own CUDA-event timers. Run it once plain (the numbers) and once under
`ncu --metrics gpu__time_duration.sum` (the launch list: which kernels the reduce stage is made of).

    python scripts/msm_small_probe.py [--logs 16,18,20] [--iters 5]

Inputs need no oracle: 1024 distinct points k*G built with Python integers, tiled to N by a seeded index map."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gnark_symmetric_crypto_b200 as G

P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _add(p, q):
    if p is None: return q
    if q is None: return p
    (x1, y1), (x2, y2) = p, q
    if x1 == x2:
        if (y1 + y2) % P == 0: return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return x3, (lam * (x1 - x3) - y1) % P


def distinct_points(m):
    """[1..m] * G as Montgomery limbs (m, 8) uint64"""
    out = np.zeros((m, 8), dtype=np.uint64)
    g = (1, 2)
    acc = None
    for i in range(m):
        acc = _add(acc, g)
        for k, v in enumerate(acc):
            vm = v * (1 << 256) % P
            for l in range(4):
                out[i, 4 * k + l] = (vm >> (64 * l)) & ((1 << 64) - 1)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logs", default="16,18,20")
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    rng = np.random.default_rng(7)
    base = distinct_points(1024)
    res = []
    for lg in [int(v) for v in args.logs.split(",")]:
        n = 1 << lg
        pts = base[rng.integers(0, len(base), n)]
        sc = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.int64).astype(np.uint64)
        sc[:, 3] &= np.uint64((1 << 60) - 1)
        for mode in ("one_shot", "fixed_base"):
            plan = G.MsmPlan(1, pts, precompute=(mode == "fixed_base"))
            plan.set_scalars(sc)
            best = None
            for _ in range(args.iters):
                out, ms = plan.run()
                if best is None or ms[0] < best[0]:
                    best = [float(v) for v in ms]
            plan.close()
            row = {"log2n": lg, "mode": mode, "ms_total": best[0], "ms_accumulate": best[1], "ms_sort": best[2], "ms_reduce": best[3],
                   "x0": int(out[0])}
            print(json.dumps(row), flush=True)
            res.append(row)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/msm_small_probe.json", "w") as f:
        json.dump(res, f)


if __name__ == "__main__":
    main()
