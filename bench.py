#!/usr/bin/env python
"""bench.py — PT scans/s of the B200 scan engine on the BASELINE.json configurations, next to the CPU
restatement of the reference path timed on the box's host cores.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on
rank 0.  The headline workload is BASELINE config 3 (8-mode Gaussian mixture d=128, AutoMALA, 1024 chains per
GPU) — the largest single-GPU configuration BASELINE.json names; configs 2, 4 and 5 at their BASELINE per-GPU
shapes ride along in the same line under `also` (each a full record with its own `roofline`, `e2e` and, at
N=1, `cpu_baseline`), so the one command the driver runs measures every named shape: with `--gpus 8` the
ladders are 8192 (C3), 2048 (C2), 4096 (C4: BASELINE's own shape) and 2048 (C5: BASELINE's own shape) chains.
`--config cX` selects another headline, `--also ""` drops the extra records.

A *step* is one `run_one_round!`-equivalent call (`pgn_run_round`) of `scans_per_step` PT scans through the C
ABI (C3: 512 scans = round 9 of a run, C2 / C4: 2048 = round 11, C5: 2 steady-state scans; rounds of a few hundred
milliseconds, so that the launch skew between the ranks of a multi-GPU run — about a millisecond — stays below 1 %).  `value` is device-timed (CUDA events around the kernels of the round, inputs resident in HBM); `e2e` is
the same metric measured around the public call with host buffers (schedule / explorer parameters copied
host->device and the round statistics copied device->host inside the timed region).

`--impl reference` times the reference's own CPU implementation of the path.  Julia is not installable
offline, so this is the oracle port (`oracle/`, built -O3 -march=native -ffp-contract=off, OpenMP over
replicas = `Threads.@threads`), run with all host threads (the count is taken from the machine, not from
OMP_NUM_THREADS, which torchrun pins to 1) on the SAME config; each step is a bounded sample of the step's
scans (stated in `cpu_baseline.sample`) so the run ends within minutes.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# BASELINE.json configs.  flops_per_point: counted FP64 operations (an fma counts 2) of ONE density + gradient
# evaluation and the leapfrog around it, per SURVEY.md §8(d): toy 2d, funnel 12d + exp, mixture 2*K*3d + K exp +
# log, + 8d for the leapfrog; exp/log counted as 30/40 operations (pgn_numerics.cuh).  Ising: integer work, 0.
CONFIGS = {
    "c1": dict(chains_per_gpu=10, dim=2, explorer="SliceSampler", state_bytes=2 * 8, scans=1024, burn=8,
               flops_per_point=2 * 2,
               kernel="pgn::scan_kernel<VecChain<TOY_MVN,1,SLICE>>",
               workload="C1: toy_mvn_target(2), SliceSampler, 10 chains (reference smoke test)"),
    "c2": dict(chains_per_gpu=256, dim=32, explorer="AutoMALA", state_bytes=32 * 8, scans=2048, burn=8,
               flops_per_point=12 * 32 + 30 + 8 * 32,
               kernel="pgn::scan_kernel<VecChain<FUNNEL,1,AUTOMALA>>",
               workload="C2: Neal's funnel d=32 (test/supporting/dimensional-analysis.jl:33-47), reference N(0,9I), "
                        "AutoMALA defaults, 256 chains per GPU, chain ladder sharded contiguously"),
    "c3": dict(chains_per_gpu=1024, dim=128, explorer="AutoMALA", state_bytes=128 * 8, scans=512, burn=8,
               flops_per_point=2 * 8 * 3 * 128 + 8 * 30 + 40 + 8 * 128,
               kernel="pgn::scan_kernel<VecChain<GMM,4,AUTOMALA,MIXED>> (blocks of two warps: the free warp slots serve the chains "
                      "that worked most as teams of two, DESIGN.md 'Mixed teams'; uniform one-warp launch with PGN_MIXED_TEAMS=0)",
               workload="C3: 8-mode Gaussian mixture d=128 (means (+-8,+-8,+-8,0,...)), reference N(0,64 I), "
                        "AutoMALA defaults, 1024 chains per GPU"),
    "c4": dict(chains_per_gpu=512, dim=1024, explorer="IsingMetropolis", state_bytes=128, scans=2048, burn=8,
               flops_per_point=0,
               kernel="pgn::scan_kernel<IsingChain>",
               workload="C4: Ising 32x32 torus (examples/ising.jl), beta = log(1+sqrt 2)/2 (critical), "
                        "IsingMetropolis(n_steps=3), 512 chains per GPU (4096 on 8 GPUs)"),
    "c5": dict(chains_per_gpu=256, dim=4096, explorer="AutoMALA", state_bytes=4096 * 8, n_data=65536, scans=2, burn=2,
               flops_per_point=4 * 65536 * 4096,
               kernel="pgn::dgemm_km_dmma_kernel (two FP64 tensor-core GEMMs per batched evaluation) + logreg_controller_kernel",
               workload="C5: synthetic logistic regression d=4096, N_data=65536 (X ~ N(0,1)/sqrt d), prior N(0,I), "
                        "AutoMALA defaults, 256 chains per GPU (2048 on 8 GPUs); steady-state scans (scan >= 2: reversed "
                        "search and MH test active)"),
}
HEADLINE = "c3"
DEFAULT_ALSO = "c2,c4,c5"


def make_target_and_explorer(pg, name):
    cfg = CONFIGS[name]
    if name == "c1":
        return pg.toy_mvn_target(2), pg.SliceSampler()
    if name == "c2":
        return pg.Funnel(32), pg.AutoMALA()
    if name == "c3":
        return pg.eight_mode_mixture(128, 8.0), pg.AutoMALA()
    if name == "c5":
        return pg.synthetic_logistic_regression(cfg["n_data"], cfg["dim"]), pg.AutoMALA()
    return pg.IsingLogPotential(0.4406867935097715, 32), pg.IsingMetropolis()


def algorithmic_bytes_per_scan(name):
    """B_scan = 2*N*state (explore: state read + written once) + N*state (swap phase re-reads each
    state) + 64*N (per-replica scalars); state = d*8 bytes, or 128 bytes for the bit-packed 32x32 Ising lattice.
    C5 adds the data term separately (see `c5_data_bytes`)."""
    cfg = CONFIGS[name]
    return 3 * cfg["chains_per_gpu"] * cfg["state_bytes"] + 64 * cfg["chains_per_gpu"]


def read_ncu_traffic(cfg_name):
    """DRAM bytes of one launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json); None when no capture of this config is committed."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            e = json.load(f).get(cfg_name)
        return e if e else None
    except Exception:
        return None


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_threads():
    """Hardware threads this process may use — NOT OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._reader, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _reader(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


RECORDER_ORDER = 0


def build_problem(pg, name, lib, n_chains, comm, device, burn_rounds, seed=1):
    """Create the PT object of config `name` and run the untimed adaptive burn-in."""
    target, explorer = make_target_and_explorer(pg, name)
    inputs = pg.Inputs(target=target, explorer=explorer, n_chains=n_chains, n_rounds=burn_rounds,
                       seed=seed, engine_lib=lib, device=device, comm=comm, recorder_order=RECORDER_ORDER)
    return pg.pigeons_pt(pg.create_pt(inputs))


def one_step(pt, scans):
    """One step through the public API: parameters host->device, the round on the
    device, statistics device->host (this is what the Julia shim does per round)."""
    eng = pt.engine
    eng.set_schedule(pt.shared.tempering.schedule.grids)
    if pt.shared.explorer is not None:
        eng.set_explorer(**pt.shared.explorer.engine_params(pt.inputs.target.dim))
    return eng.run_round(scans)


BASELINE_LADDER = {"c1": 10, "c2": 256, "c3": 1024, "c4": 4096, "c5": 2048}     # n_chains BASELINE.json names


def ladder_chains(name, gpus, scaling):
    """weak: chains_per_gpu per GPU; strong: the ladder BASELINE.json names, split over the GPUs"""
    return BASELINE_LADDER[name] if scaling == "strong" else CONFIGS[name]["chains_per_gpu"] * gpus


def config_dict(name, gpus, scans, burn, scaling="weak"):
    cfg = CONFIGS[name]
    cpg = cfg["chains_per_gpu"]
    if scaling == "strong":
        n = BASELINE_LADDER[name]
        return {"workload": cfg["workload"], "name": name, "n_chains": n, "dim": cfg["dim"], "explorer": cfg["explorer"],
                "scans_per_step": scans, "burn_in_rounds": burn, "parallelism": f"chains/{gpus}",
                "scan_unit": f"one PT scan of the {n}-chain ladder BASELINE.json names, split contiguously over the N GPUs (strong scaling)",
                "l2": "flushed between timed steps (256 MiB write); the working set is register-resident"}
    return {"workload": cfg["workload"], "name": name, "n_chains": cpg * gpus, "dim": cfg["dim"], "explorer": cfg["explorer"],
            "scans_per_step": scans, "burn_in_rounds": burn, "parallelism": f"chains/{gpus}",
            "scan_unit": f"one PT scan of {cpg} chains; with N GPUs the ladder has {cpg}*N chains and value = N * ladder scans/s",
            "l2": "flushed between timed steps (256 MiB write); the working set is register-resident"}


# ---------------------------------------------------------------------------------------------- CPU (oracle) legs
def cpu_oracle_problem(pg, name, n_chains, burn_rounds, clone_from=None, burn_budget_s=30.0):
    """The oracle port holding the same workload.  `clone_from`: a PT whose adapted schedule / explorer / replica
    states are copied instead of re-running the burn-in on the CPU (same workload state as the GPU arm).  Without it
    the adaptive burn-in runs on the CPU with every host thread, bounded to the rounds that fit `burn_budget_s`
    (the burn-in only brings the schedule and the step size to a representative state; it is not timed)."""
    from oracle_adapter import load_oracle
    import ctypes as C
    lib = load_oracle()
    target, explorer = make_target_and_explorer(pg, name)
    pt = pg.create_pt(pg.Inputs(target=target, explorer=explorer, n_chains=n_chains, n_rounds=0, seed=1, engine_lib=lib))
    lib.lib.orc_set_threads(pt.engine._h, C.c_int(host_threads()))       # not omp_get_max_threads(): torchrun sets OMP_NUM_THREADS=1
    burned = 0
    if clone_from is None:
        t0 = time.perf_counter()
        while burned < burn_rounds:
            pt.inputs.n_rounds = burned + 1
            pt = pg.pigeons_pt(pt)
            burned += 1
            spent = time.perf_counter() - t0
            if spent * 2.0 > burn_budget_s - spent:      # the next round costs twice this one
                break
    else:
        pt.shared = clone_from.shared
        st = clone_from.engine.get_state()
        pt.engine.set_state(x=st["x"] if st["x"].size else None, replica_index=st["replica_index"],
                            rng_counter=st["rng_counter"], round_trip_state=st["round_trip_state"])
    pt.cpu_burn_rounds = burned
    return pt, lib, C


def pick_threads(pt, lib, C, threads):
    """OpenMP over replicas does not always scale to every hardware thread; the reference arm gets its best
    configuration from a short sweep (threads=0) over {all, 1/2, 1/4 of the hardware threads, 16, 8}."""
    hw = host_threads()
    cands = [min(threads, hw)] if threads else sorted({t for t in (hw, hw // 2, hw // 4, 16, 8) if 1 <= t <= hw})
    best = None
    for t in cands:
        lib.lib.orc_set_threads(pt.engine._h, C.c_int(t))
        one_step(pt, 2)
        r = one_step(pt, 4)
        per_scan = max(r.wall_s / 4, 1e-9)
        if best is None or per_scan < best[1]:
            best = (t, per_scan)
    lib.lib.orc_set_threads(pt.engine._h, C.c_int(best[0]))
    return best


def cpu_baseline_scan(pg, name, threads, budget_s, clone_from):
    cfg = CONFIGS[name]
    cpt, clib, C = cpu_oracle_problem(pg, name, cfg["chains_per_gpu"], 0, clone_from=clone_from)
    n_threads, per_scan = pick_threads(cpt, clib, C, threads)
    scans = int(max(8, min(cfg["scans"], budget_s / per_scan)))
    r = one_step(cpt, scans)
    cpt.close()
    return {"value": scans / r.wall_s, "unit": "scans/s", "cores": n_threads, "kind": "port", "host_threads": host_threads(),
            "sample": f"{scans} scans of the same {cfg['chains_per_gpu']}-chain workload, continuing from the GPU arm's adapted "
                      "state (oracle/: C++ restatement of the reference path, -O3 -march=native, OpenMP over replicas)"}


def cpu_baseline_c5(pg, threads=0):
    """C5 on the CPU is far too slow to time whole scans (one density+gradient evaluation streams the 2 GiB design
    matrix twice per chain).  Bounded sample: T chains (one per host thread) x ONE momentum refreshment of the same
    autoMALA kernel on the same data, scans 2..3 (MH and reversed search active), extrapolated linearly to 256 chains x
    57 refreshments (explore is independent per replica and per refreshment)."""
    from oracle_adapter import load_oracle
    import ctypes as C
    cfg = CONFIGS["c5"]
    lib = load_oracle()
    target = pg.synthetic_logistic_regression(cfg["n_data"], cfg["dim"])
    t_chains = min(threads or host_threads(), 64)
    explorer = pg.AutoMALA(base_n_refresh=1, exponent_n_refresh=0.0, step_size=0.02)
    pt = pg.create_pt(pg.Inputs(target=target, explorer=explorer, n_chains=t_chains, n_rounds=0, seed=1, engine_lib=lib))
    lib.lib.orc_set_threads(pt.engine._h, C.c_int(t_chains))
    pt.engine.set_schedule(pt.shared.tempering.schedule.grids)
    pt.engine.set_explorer(**explorer.engine_params(cfg["dim"]))
    r = pt.engine.run_round(3)          # scan 1 has no MH step; scans 2, 3 are steady-state scans
    full_refresh = pg.AutoMALA().n_refresh(cfg["dim"])
    per_chain_refresh_s = r.wall_s / 3.0          # T chains in parallel on T threads, 1 refreshment per scan
    scan_s = per_chain_refresh_s * full_refresh * (cfg["chains_per_gpu"] / t_chains)
    pt.close()
    return {"value": 1.0 / scan_s, "unit": "scans/s", "cores": t_chains, "kind": "port", "host_threads": host_threads(),
            "sample": f"{t_chains} chains x 3 scans x 1 refreshment of the same kernel and data on {t_chains} threads "
                      f"({r.wall_s:.1f} s), extrapolated x{full_refresh} refreshments x{cfg['chains_per_gpu']}/{t_chains} chains"}


def reference_record(pg, name, args, budget_s):
    """`--impl reference` for one config: the oracle port on the host cores, same config dict as the GPU arm."""
    cfg = CONFIGS[name]
    gpus = max(args.gpus, 1)
    n_chains = ladder_chains(name, gpus, args.scaling)
    scans_cfg = args.scans if (args.scans and name == args.config) else cfg["scans"]
    burn = args.burn_rounds if (args.burn_rounds is not None and name == args.config) else cfg["burn"]
    config = config_dict(name, gpus, scans_cfg, burn, args.scaling)
    if args.scaling == "strong":
        gpus = 1          # value = scans/s of the one ladder
    if name == "c5":
        b = cpu_baseline_c5(pg, args.cpu_threads)
        value = b["value"] / gpus * gpus      # ladder rate x N GPUs' worth of chains = the same extrapolation
        return {"impl": "reference", "metric": "pt_scans_per_s", "value": value, "unit": "scans/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * scans_cfg / b["value"], "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": b, "e2e": {"value": value, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    pt, lib, C = cpu_oracle_problem(pg, name, n_chains, burn)
    n_threads, per_scan = pick_threads(pt, lib, C, args.cpu_threads)
    total_steps = max(args.steps, 1) + max(args.warmup, 0)
    sample = int(max(4, min(scans_cfg, budget_s / total_steps / per_scan)))
    for _ in range(max(args.warmup, 0)):
        one_step(pt, sample)
    t = 0.0
    for _ in range(args.steps):
        t += one_step(pt, sample).wall_s
    burned = pt.cpu_burn_rounds
    pt.close()
    ladder = args.steps * sample / t
    value = ladder * gpus
    return {"impl": "reference", "metric": "pt_scans_per_s", "value": value, "unit": "scans/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * scans_cfg / ladder, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": "scans/s", "cores": n_threads, "kind": "port", "host_threads": host_threads(),
                             "sample": f"{args.steps} timed steps, each a bounded sample of {sample} of the step's {scans_cfg} "
                                       f"scans of the {n_chains}-chain ladder (ms_per_step is scaled to the full step), after "
                                       f"{burned} adaptive burn-in rounds on the CPU; oracle/ "
                                       "(C++ restatement of the reference path, -O3 -march=native, OpenMP over replicas)"},
            "e2e": {"value": value, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "Julia/Pigeons.jl is not installable offline; the reference arm is the CPU restatement in oracle/"}


# ---------------------------------------------------------------------------------------------- B200 arm
def b200_record(pg, torch, name, args, lib, comm, rank, world, local_rank, steps, warmup, flush, with_cpu):
    import numpy as np
    cfg = CONFIGS[name]
    gpus = max(args.gpus, 1)
    cpg = cfg["chains_per_gpu"]
    strong = args.scaling == "strong"
    n_chains = ladder_chains(name, gpus, args.scaling)
    scans = args.scans if (args.scans and name == args.config) else cfg["scans"]
    burn = args.burn_rounds if (args.burn_rounds is not None and name == args.config) else cfg["burn"]
    pt = build_problem(pg, name, lib, n_chains, comm, local_rank, burn)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            comm.barrier()
            torch.cuda.synchronize()

    for _ in range(max(warmup, 0)):
        one_step(pt, scans)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sync_all()
    kernel_ms = wall_s = gemm_ms = 0.0
    batch_steps = pts = evals = launches = act_cols = gemm_cols = 0
    t_region0 = time.perf_counter()
    for _ in range(steps):
        flush.fill_(1)                       # L2 flush, outside the timed intervals
        sync_all()
        t0 = time.perf_counter()
        r = one_step(pt, scans)              # synchronous: returns after the stream sync + D2H of the statistics
        wall_s += time.perf_counter() - t0
        kernel_ms += r.kernel_ms
        gemm_ms += r.gemm_ms
        batch_steps += r.batch_steps
        pts += r.n_density_points
        evals += r.n_ref_equiv_evals
        launches += r.n_launches
        act_cols += r.active_columns
        gemm_cols += r.gemm_columns
    sync_all()
    t_region = time.perf_counter() - t_region0
    clocks = sampler.stop() if rank == 0 else None

    # max over ranks of the device time / wall time; sums of the work counters
    pts_local = pts
    if world > 1:
        tt = np.stack(comm.all_gather_array(np.array([kernel_ms, wall_s])))
        kernel_ms, wall_s = float(tt[:, 0].max()), float(tt[:, 1].max())
        cc = np.stack(comm.all_gather_array(np.array([pts, evals, launches], dtype=np.int64))).sum(axis=0)
        pts, evals, launches = int(cc[0]), int(cc[1]), int(cc[2])
    n_local = pt.engine.n_local
    line = None
    if rank == 0:
        total_scans = steps * scans
        ladder_scans_per_s = total_scans / (kernel_ms * 1e-3)
        value = ladder_scans_per_s * (1 if strong else gpus)
        e2e_value = total_scans / wall_s * (1 if strong else gpus)
        peak, peak_src = read_peaks()
        launch_ms = kernel_ms / steps
        b_scan = algorithmic_bytes_per_scan(name) * (pt.engine.n_local / cpg)      # this GPU's chains
        data_bytes = 0.0
        if name == "c5" and batch_steps > 0:
            # SURVEY §8(d) data term P*|D|: every batched evaluation streams X (2 GiB) once per GEMM
            data_bytes = 2.0 * cfg["n_data"] * cfg["dim"] * 8 * batch_steps / steps
        achieved = (b_scan * scans + data_bytes) / (launch_ms * 1e-3) / 1e9
        tr = read_ncu_traffic(name)
        fp64_peak = lib.measure_fp64_peak(local_rank)
        flops = float(cfg["flops_per_point"]) * pts_local            # this GPU's counted operations in the timed region
        fp64_ach = flops / (kernel_ms * 1e-3) / 1e12
        latency_bound = name != "c5"
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": tr["dram_bytes_per_launch"] if tr else None,
                "traffic_note": ((f"DRAM bytes of one {tr['scans_in_captured_launch']}-scan launch, {tr['source']}; the replica state "
                                  "stays on chip for the whole launch, so the traffic does not grow with the number of scans")
                                 if tr.get("scans_in_captured_launch") else f"DRAM bytes of one launch, {tr['source']}") if tr else None,
                "peak_source": peak_src, "kernel": cfg["kernel"] + (" (one persistent launch per step)" if latency_bound else ""),
                "algorithmic_bytes_per_scan": b_scan,
                "fp64": {"achieved": fp64_ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp64_ach / fp64_peak if fp64_peak else None,
                         "flops_per_density_point": cfg["flops_per_point"], "density_points": pts_local,
                         "peak_source": "pgn_measure_fp64_peak in this run (register-resident DFMA loop on all SMs)",
                         "pipe_active_pct_ncu": tr.get("fp64_pipe_active_pct") if tr else None},
                "binding_bound": ("FP64 dependency latency along each replica's serial chain (the working set is on chip for the whole "
                                  "round; neither roof is approached by design, SURVEY.md §8d)") if latency_bound else
                                 "FP64 tensor pipe (DMMA): compute-bound dense contraction"}
        line = {
            "metric": "pt_scans_per_s", "value": value, "unit": "scans/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": launch_ms, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(name, gpus, scans, burn, args.scaling),
            "ladder_scans_per_s": ladder_scans_per_s,
            "log_potential_evals_per_s": {"ref_equiv": evals / (kernel_ms * 1e-3), "unique_points": pts / (kernel_ms * 1e-3),
                                          "note": "ref_equiv = log_potential/logdensity[_and_gradient] calls the reference code "
                                                  "path makes for the same trajectory; unique = density points the kernel evaluates"},
            "e2e": {"value": e2e_value, "unit": "scans/s",
                    "h2d_bytes_per_step": n_chains * 8 + 184 + (cfg["dim"] * 8 if cfg["explorer"] == "AutoMALA" else 0),
                    "d2h_bytes_per_step": n_local * 160 + 2 * cfg["dim"] * 8 + 12,
                    "note": "wall clock around set_schedule + set_explorer + pgn_run_round (host buffers in, statistics out)"},
            "gpu_launches": launches,
            "roofline": roof,
            "clocks": clocks,
            "timed_region_s": t_region,
        }
        if name == "c5" and batch_steps > 0:
            flops_per_col = 2 * 2.0 * cfg["n_data"] * cfg["dim"]                      # two GEMMs, 2 flops per fma, one column
            gemm_tflops = flops_per_col * gemm_cols / (gemm_ms * 1e-3) / 1e12
            line["fp64_roofline"] = {
                "bound": "fp64 (tcgen05 has no FP64 path; DMMA m8n8k4 GEMM whose summation order equals the sequential-fma spec; "
                         "PGN_GEMM=simt selects the DFMA kernel)",
                "gemm": os.environ.get("PGN_GEMM", "dmma"),
                "achieved_tflops": gemm_tflops, "peak_tflops": fp64_peak,
                "frac": gemm_tflops / fp64_peak if fp64_peak else None,
                "dmma_pipe_active_pct_ncu": (tr or {}).get("dmma_pipe_active_pct"),
                "gemm_share_of_kernel_time": gemm_ms / kernel_ms, "batched_evaluations": batch_steps,
                "columns_requested": act_cols, "columns_multiplied": gemm_cols,
                "useful_flop_fraction": act_cols / gemm_cols if gemm_cols else None,
                "column_utilisation_of_256": act_cols / (batch_steps * cpg) if batch_steps else None,
                "flops_per_column_evaluation": flops_per_col}
    if with_cpu and rank == 0 and gpus == 1:
        line["cpu_baseline"] = cpu_baseline_c5(pg, args.cpu_threads) if name == "c5" else \
            cpu_baseline_scan(pg, name, args.cpu_threads, 10.0, pt)
    pt.close()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scans", type=int, default=None, help="PT scans per step of the headline config (default: its table entry)")
    ap.add_argument("--burn-rounds", type=int, default=None, help="adaptive burn-in rounds before timing (headline config)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default=HEADLINE, choices=sorted(CONFIGS), help=f"headline BASELINE.json config (default {HEADLINE})")
    ap.add_argument("--also", default=None, help=f"configs carried as extra records in the same line (default '{DEFAULT_ALSO}' "
                                                 "when --config is not given, else none)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): chains per GPU fixed, the ladder grows with N; strong: the BASELINE ladder of the config "
                         "(chains_per_gpu chains in total) is split over the N GPUs")
    ap.add_argument("--recorder-order", type=int, default=0, choices=[0, 1],
                    help="pgn_config.recorder_order: 0 per-replica recorders + tree merge (reference), 1 per chain")
    ap.add_argument("--cpu-threads", type=int, default=0, help="threads of the CPU baseline (0 = pick the fastest of a sweep)")
    args = ap.parse_args()
    explicit = any(a == "--config" or a.startswith("--config=") for a in sys.argv[1:])
    also = args.also if args.also is not None else ("" if explicit else DEFAULT_ALSO)
    also = [c for c in also.split(",") if c and c != args.config]
    for c in also:
        if c not in CONFIGS:
            ap.error(f"unknown config in --also: {c}")

    global RECORDER_ORDER
    RECORDER_ORDER = args.recorder_order
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import pigeons_jl_b200 as pg

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        line = reference_record(pg, args.config, args, budget_s=100.0 if args.config != "c5" else 0)
        extra = {}
        for c in also:
            if c in ("c4", "c5"):
                continue      # carried by the B200 arm's cpu_baseline; the reference arm stays within minutes
            extra[c] = reference_record(pg, c, args, budget_s=40.0)
        if extra:
            line["also"] = extra
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    torch.cuda.set_device(local_rank)
    comm = pg.SingleProcess()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = pg.TorchDistributed(device=torch.device("cuda", local_rank))
    lib = pg.EngineLib()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    with_cpu = not args.no_cpu_baseline
    line = b200_record(pg, torch, args.config, args, lib, comm, rank, world, local_rank, args.steps, args.warmup, flush, with_cpu)
    extra = {}
    for c in also:
        st, wu = (2, 1) if c == "c5" else (args.steps, args.warmup)
        rec = b200_record(pg, torch, c, args, lib, comm, rank, world, local_rank, st, wu, flush, with_cpu)
        if rank == 0:
            extra[c] = rec
    if rank == 0:
        if extra:
            line["also"] = extra
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
