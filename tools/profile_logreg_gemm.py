"""Drives one batched density+gradient evaluation of the C5 logistic-regression target
(two FP64 GEMMs) through the C ABI, for `ncu -k regex:dgemm_km`.

    ncu --set full --clock-control none --import-source on -k regex:dgemm_km -c 2 \
        -o gpurun_out/prof_gemm python tools/profile_logreg_gemm.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                 # noqa: E402
import pigeons_jl_b200 as pg       # noqa: E402

n_data = int(os.environ.get("N_DATA", 65536))
dim = int(os.environ.get("DIM", 4096))
chains = int(os.environ.get("CHAINS", 256))
t = pg.synthetic_logistic_regression(n_data, dim)
e = pg.Engine(pg.EngineLib(), n_chains=chains, seed=1, **t.engine_config())
x = np.random.default_rng(0).normal(0, 0.5, (chains, dim))
beta = np.linspace(0, 1, chains)
reps = int(os.environ.get("REPS", 1))
for _ in range(reps):
    t0 = time.perf_counter()
    ld, g = e.logdensity_and_gradient(x, beta)
    dt = time.perf_counter() - t0
print("batched evaluation wall", dt, "s; logdens[:3]", ld[:3])
