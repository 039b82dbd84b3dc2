"""Extra measurement (not the bench.py contract): BASELINE config 3's ladder run as TWO legs of 512 chains
(StabilizedPT), with and without the Gaussian variational reference, next to the single-leg ladder.
Device-timed scans/s of 512-scan rounds after an 8-round adaptive burn-in.

    python tools/bench_two_legs.py > gpurun_out/two_legs.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np            # noqa: E402

import pigeons_jl_b200 as pg  # noqa: E402


def measure(label, steps=3, scans=512, burn=8, **kw):
    pt = pg.pigeons(target=pg.eight_mode_mixture(128, 8.0), explorer=pg.AutoMALA(), n_rounds=burn, seed=1, **kw)
    eng = pt.engine
    eng.set_schedule(pg.tempering_parameters(pt.shared.tempering))
    var = pt.inputs.variational
    if pt.inputs.n_chains_variational > 0:
        eng.set_variational(*( (var.mean, var.standard_deviation) if var is not None and var.mean is not None else (None, None)))
    eng.set_explorer(**pt.shared.explorer.engine_params(128))
    ms = []
    for _ in range(steps + 1):
        ms.append(eng.run_round(scans).kernel_ms)
    ms = ms[1:]
    out = dict(config=label, n_chains=pt.inputs.n_chains_total, scans_per_step=scans, ms_per_step=float(np.mean(ms)),
               scans_per_s=scans / (float(np.mean(ms)) * 1e-3), global_barrier=pg.global_barrier(pt))
    if isinstance(pt.shared.tempering, pg.StabilizedPT):
        out["global_barrier_variational"] = pg.global_barrier_variational(pt)
    pt.close()
    return out


if __name__ == "__main__":
    res = [measure("one leg, 1024 chains (C3)", n_chains=1024),
           measure("two legs 512 + 512, fixed reference on both", n_chains=512, n_chains_variational=512),
           measure("two legs 512 + 512, GaussianReference from round 3", n_chains=512, n_chains_variational=512,
                   variational=pg.GaussianReference(first_tuning_round=3))]
    print(json.dumps(res, indent=1))
