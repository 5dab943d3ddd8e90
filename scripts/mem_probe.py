import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
import gnark_symmetric_crypto_b200 as G
free0, total = torch.cuda.mem_get_info(0)
ctx = G.Groth16Context(open("tests/golden/pk.chacha20", "rb").read(), open("tests/golden/r1cs.chacha20", "rb").read())
free1, _ = torch.cuda.mem_get_info(0)
k, no, c, i, r = ctx._pack(*bench.make_requests(1024, b"g16-b200-batch"))[1:]
ctx.stage(k, no, c, i, r)
for _ in range(3): ctx.run()
free2, _ = torch.cuda.mem_get_info(0)
print("after init GB", (free0 - free1) / 1e9, "after 3 batches of 1024 GB", (free0 - free2) / 1e9, "total", total / 1e9)
