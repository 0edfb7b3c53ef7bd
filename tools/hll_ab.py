import sys, json, numpy as np
sys.path.insert(0, ".")
import torch
from lash_b200 import ALGO_HLL, ops
from tools import bench_configs as B
rng = np.random.default_rng(3)
p = 14; m = 1 << p
rho = np.clip(np.floor(8.0 - np.log2(-np.log(rng.random((3000, m))))) + 1, 1, 51).astype(np.uint8)
with ops.Context(0) as ctx:
    for n in (1000, 3000):
        print(json.dumps(B.dist_case(ctx, f"HLL p=14 n={n}", ALGO_HLL, 14, 21, 0, rho[:n])), flush=True)
    d, _ = ops.dist(ctx, ALGO_HLL, 14, 21, 0, 1, False, rho[:300], rho[:300])
    np.save("gpurun_out/hll_d_%s.npy" % sys.argv[1], d)
