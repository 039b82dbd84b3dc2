#!/usr/bin/env python
"""bench.py — PT scans/s of the B200 scan engine on BASELINE config 2
(Neal's funnel d=32, AutoMALA, 256 chains per GPU), next to the CPU restatement
of the reference path timed on the box's host cores.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W`
prints ONE JSON line on rank 0.  A *step* is one `run_one_round!`-equivalent call
of `--scans` PT scans (default 1024 = the last round of a 10-round run) through
the C ABI.  `value` is device-timed (CUDA events around the scan kernel, inputs
resident in HBM); `e2e` is the same metric measured around the public call with
host buffers (schedule / explorer parameters copied host->device and the round
statistics copied device->host inside the timed region).

`--impl reference` times the reference's own CPU implementation of the path.
Julia is not installable here, so this is the oracle port (`oracle/`), run with
all host threads, on bounded samples of the same workload.
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

BURN_IN_ROUNDS = 8          # rounds 1..8 (510 scans) with adaptation, untimed set-up

# BASELINE.json configs.  The driver runs the default (c2 = the configuration the metric is quoted
# on that fits one GPU); the others are run by hand and their lines are committed under profiles/.
CONFIGS = {
    "c1": dict(chains_per_gpu=10, dim=2, explorer="SliceSampler", state_bytes=2 * 8,
               kernel="pgn::scan_kernel<VecChain<TOY_MVN,1,SLICE>>",
               workload="C1: toy_mvn_target(2), SliceSampler, 10 chains (reference smoke test)"),
    "c2": dict(chains_per_gpu=256, dim=32, explorer="AutoMALA", state_bytes=32 * 8,
               kernel="pgn::scan_kernel<VecChain<FUNNEL,1,AUTOMALA>>",
               workload="C2: Neal's funnel d=32 (test/supporting/dimensional-analysis.jl:33-47), reference N(0,9I), "
                        "AutoMALA defaults, 256 chains per GPU, chain ladder sharded contiguously"),
    "c3": dict(chains_per_gpu=1024, dim=128, explorer="AutoMALA", state_bytes=128 * 8,
               kernel="pgn::scan_kernel<VecChain<GMM,4,AUTOMALA>>",
               workload="C3: 8-mode Gaussian mixture d=128 (means (+-8,+-8,+-8,0,...)), reference N(0,64 I), "
                        "AutoMALA defaults, 1024 chains per GPU"),
    "c4": dict(chains_per_gpu=512, dim=1024, explorer="IsingMetropolis", state_bytes=128,
               kernel="pgn::scan_kernel<IsingChain>",
               workload="C4: Ising 32x32 torus (examples/ising.jl), beta = log(1+sqrt 2)/2 (critical), "
                        "IsingMetropolis(n_steps=3), 512 chains per GPU (4096 on 8 GPUs)"),
    "c5": dict(chains_per_gpu=256, dim=4096, explorer="AutoMALA", state_bytes=4096 * 8, n_data=65536,
               kernel="pgn::dgemm_km_dmma_kernel<0/1> (two FP64 tensor-core GEMMs per batched evaluation) + logreg_controller_kernel",
               workload="C5: synthetic logistic regression d=4096, N_data=65536 (X ~ N(0,1)/sqrt d), prior N(0,I), "
                        "AutoMALA defaults, 256 chains per GPU"),
}
CFG = CONFIGS["c2"]
CHAINS_PER_GPU = CFG["chains_per_gpu"]
DIM = CFG["dim"]


def select_config(name):
    global CFG, CHAINS_PER_GPU, DIM
    CFG = CONFIGS[name]
    CHAINS_PER_GPU = CFG["chains_per_gpu"]
    DIM = CFG["dim"]


def make_target_and_explorer(pg, name):
    if name == "c1":
        return pg.toy_mvn_target(2), pg.SliceSampler()
    if name == "c2":
        return pg.Funnel(32), pg.AutoMALA()
    if name == "c3":
        return pg.eight_mode_mixture(128, 8.0), pg.AutoMALA()
    if name == "c5":
        return pg.synthetic_logistic_regression(CFG["n_data"], CFG["dim"]), pg.AutoMALA()
    return pg.IsingLogPotential(0.4406867935097715, 32), pg.IsingMetropolis()


def algorithmic_bytes_per_scan(n_chains, dim=None):
    """B_scan = 2*N*state (explore: state read + written once)
    + N*state (swap phase re-reads each state) + 64*N (per-replica scalars);
    state = d*8 bytes, or 128 bytes for the bit-packed 32x32 Ising lattice."""
    return 3 * n_chains * CFG["state_bytes"] + 64 * n_chains


def read_ncu_traffic(cfg_name):
    """DRAM bytes of one launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json); None when no capture of this config is committed."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            e = json.load(f).get(cfg_name)
        return (e["dram_bytes_per_launch"], e["scans_in_captured_launch"], e["source"]) if e else (None, None, None)
    except Exception:
        return None, None, None


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


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


CFG_NAME = "c2"
BURN_ROUNDS_OVERRIDE = None


def build_problem(pg, lib, n_chains, comm, device, seed=1):
    """C2 workload: create the PT object and run the untimed adaptive burn-in."""
    target, explorer = make_target_and_explorer(pg, CFG_NAME)
    inputs = pg.Inputs(target=target, explorer=explorer, n_chains=n_chains,
                       n_rounds=BURN_ROUNDS_OVERRIDE if BURN_ROUNDS_OVERRIDE is not None else BURN_IN_ROUNDS,
                       seed=seed, engine_lib=lib, device=device, comm=comm)
    pt = pg.create_pt(inputs)
    pt = pg.pigeons_pt(pt)
    return pt


def one_step(pt, pg, scans):
    """One step through the public API: parameters host->device, the round on the
    device, statistics device->host (this is what the Julia shim does per round)."""
    eng = pt.engine
    eng.set_schedule(pt.shared.tempering.schedule.grids)
    if pt.shared.explorer is not None:
        eng.set_explorer(**pt.shared.explorer.engine_params(DIM))
    return eng.run_round(scans)


def cpu_reference_run(pg, scans_hint, n_chains, threads=0, budget_s=12.0, clone_from=None):
    """Oracle port on the host cores: same workload, bounded sample.  With threads=0 a short
    sweep over thread counts picks the fastest one (OpenMP over replicas does not always scale to
    every hardware thread; the reference arm should get its best configuration).  `clone_from`:
    a PT whose adapted schedule / explorer / replica states are copied instead of re-running the
    burn-in on the CPU (same workload state as the GPU arm)."""
    from oracle_adapter import load_oracle
    import ctypes as C
    lib = load_oracle()
    if clone_from is None:
        pt = build_problem(pg, lib, n_chains, pg.SingleProcess(), 0)
    else:
        target, explorer = make_target_and_explorer(pg, CFG_NAME)
        pt = pg.create_pt(pg.Inputs(target=target, explorer=explorer, n_chains=n_chains, n_rounds=0, seed=1, engine_lib=lib))
        pt.shared = clone_from.shared
        st = clone_from.engine.get_state()
        pt.engine.set_state(x=st["x"] if st["x"].size else None, replica_index=st["replica_index"],
                            rng_counter=st["rng_counter"], round_trip_state=st["round_trip_state"])
    max_threads = lib.lib.orc_get_threads(pt.engine._h)
    probe = 8
    if threads:
        candidates = [min(threads, max_threads)]
    else:
        candidates = sorted({t for t in (max_threads, max_threads // 2, max_threads // 4, 16, 8) if 1 <= t <= max_threads})
    best = None
    for t in candidates:
        lib.lib.orc_set_threads(pt.engine._h, C.c_int(t))
        one_step(pt, pg, 2)
        r = one_step(pt, pg, probe)
        per_scan = max(r.wall_s / probe, 1e-9)
        if best is None or per_scan < best[1]:
            best = (t, per_scan)
    n_threads, per_scan = best
    lib.lib.orc_set_threads(pt.engine._h, C.c_int(n_threads))
    scans = int(max(16, min(scans_hint, budget_s / per_scan)))
    return pt, lib, n_threads, scans


def cpu_baseline_c5(pg, threads=0):
    """C5 on the CPU is far too slow to time whole scans (one density+gradient evaluation streams
    the 2 GiB design matrix twice per chain).  Bounded sample: T chains (one per host thread) x ONE
    momentum refreshment of the same autoMALA kernel on the same data, extrapolated linearly to
    256 chains x 57 refreshments (explore is independent per replica and per refreshment)."""
    from oracle_adapter import load_oracle
    import ctypes as C
    lib = load_oracle()
    target = pg.synthetic_logistic_regression(CFG["n_data"], CFG["dim"])
    probe = pg.create_pt(pg.Inputs(target=target, explorer=pg.AutoMALA(), n_chains=2, n_rounds=0, seed=1, engine_lib=lib))
    max_threads = lib.lib.orc_get_threads(probe.engine._h)
    probe.close()
    t_chains = threads or max_threads
    explorer = pg.AutoMALA(base_n_refresh=1, exponent_n_refresh=0.0, step_size=0.02)
    pt = pg.create_pt(pg.Inputs(target=target, explorer=explorer, n_chains=t_chains, n_rounds=0, seed=1, engine_lib=lib))
    lib.lib.orc_set_threads(pt.engine._h, C.c_int(t_chains))
    pt.engine.set_schedule(pt.shared.tempering.schedule.grids)
    pt.engine.set_explorer(**explorer.engine_params(CFG["dim"]))
    r = pt.engine.run_round(2)          # scan 1 has no MH step; time both, count refreshments actually done
    full_refresh = pg.AutoMALA().n_refresh(CFG["dim"])
    per_chain_refresh_s = r.wall_s / 2.0          # T chains in parallel on T threads, 1 refreshment per scan
    scan_s = per_chain_refresh_s * full_refresh * (CHAINS_PER_GPU / t_chains)
    return {"value": 1.0 / scan_s, "unit": "scans/s", "cores": t_chains, "kind": "port",
            "sample": f"{t_chains} chains x 2 scans x 1 refreshment of the same kernel and data on {t_chains} threads "
                      f"({r.wall_s:.1f} s), extrapolated x{full_refresh} refreshments x{CHAINS_PER_GPU}/{t_chains} chains"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scans", type=int, default=None, help="PT scans per step (one run_round call); default 1024 (c5: 1)")
    ap.add_argument("--burn-rounds", type=int, default=None, help="adaptive burn-in rounds before timing (default 8; c5: 2)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS), help="BASELINE.json config (default c2)")
    ap.add_argument("--cpu-threads", type=int, default=0, help="threads of the CPU baseline (0 = pick the fastest of a sweep)")
    args = ap.parse_args()
    global CFG_NAME, BURN_ROUNDS_OVERRIDE
    CFG_NAME = args.config
    select_config(args.config)
    if args.burn_rounds is not None:
        BURN_ROUNDS_OVERRIDE = args.burn_rounds
    elif args.config == "c5":
        BURN_ROUNDS_OVERRIDE = 2          # a C5 scan is seconds of FP64 GEMMs
    if args.scans is None:
        args.scans = 1 if args.config == "c5" else 1024

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import pigeons_jl_b200 as pg

    n_chains = CHAINS_PER_GPU * max(args.gpus, 1)
    config = {"workload": CFG["workload"],
              "n_chains": n_chains, "dim": DIM, "explorer": CFG["explorer"], "scans_per_step": args.scans,
              "burn_in_rounds": BURN_ROUNDS_OVERRIDE if BURN_ROUNDS_OVERRIDE is not None else BURN_IN_ROUNDS, "parallelism": f"chains/{args.gpus}",
              "scan_unit": f"one PT scan of {CHAINS_PER_GPU} chains; with N GPUs the ladder has {CHAINS_PER_GPU}*N chains "
                           "and value = N * ladder scans/s",
              "l2": "flushed between timed steps (256 MiB write); the working set is register-resident"}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        pt, lib, n_threads, scans = cpu_reference_run(pg, args.scans, n_chains, threads=args.cpu_threads, budget_s=10.0)
        for _ in range(max(args.warmup, 0)):
            one_step(pt, pg, min(scans, 16))
        t = 0.0
        for _ in range(args.steps):
            t += one_step(pt, pg, scans).wall_s
        ladder = args.steps * scans / t
        value = ladder * max(args.gpus, 1)
        config["scans_per_step"] = scans
        line = {"impl": "reference", "metric": "pt_scans_per_s", "value": value, "unit": "scans/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": value, "unit": "scans/s", "cores": n_threads, "kind": "port",
                                 "sample": f"{args.steps} steps x {scans} scans of the {n_chains}-chain ladder, "
                                           "oracle/ (C++ restatement of the reference path, OpenMP over replicas)"},
                "e2e": {"value": value, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "Julia/Pigeons.jl is not installable offline; the reference arm is the CPU restatement in oracle/"}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ B200 arm
    os.environ["NCCL_DEBUG"] = os.environ.get("PGN_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
    import torch
    torch.cuda.set_device(local_rank)
    comm = pg.SingleProcess()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = pg.TorchDistributed(device=torch.device("cuda", local_rank))
    lib = pg.EngineLib()
    pt = build_problem(pg, lib, n_chains, comm, local_rank)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            comm.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        one_step(pt, pg, args.scans)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sync_all()
    kernel_ms, wall_s = 0.0, 0.0
    gemm_ms, batch_steps = 0.0, 0
    pts = evals = 0
    t_region0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)                       # L2 flush, outside the timed intervals
        sync_all()
        t0 = time.perf_counter()
        r = one_step(pt, pg, args.scans)     # synchronous: returns after the stream sync + D2H of the statistics
        wall_s += time.perf_counter() - t0
        kernel_ms += r.kernel_ms
        gemm_ms += r.gemm_ms
        batch_steps += r.batch_steps
        pts += r.n_density_points
        evals += r.n_ref_equiv_evals
    sync_all()
    t_region = time.perf_counter() - t_region0
    clocks = sampler.stop() if rank == 0 else None

    # max over ranks of the device time / wall time; sums of the work counters
    import numpy as np
    if world > 1:
        tt = np.stack(comm.all_gather_array(np.array([kernel_ms, wall_s])))
        kernel_ms, wall_s = float(tt[:, 0].max()), float(tt[:, 1].max())
        cc = np.stack(comm.all_gather_array(np.array([pts, evals], dtype=np.int64))).sum(axis=0)
        pts, evals = int(cc[0]), int(cc[1])
    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return

    total_scans = args.steps * args.scans
    ladder_scans_per_s = total_scans / (kernel_ms * 1e-3)
    value = ladder_scans_per_s * args.gpus
    e2e_value = total_scans / wall_s * args.gpus
    peak, peak_src = read_peaks()
    bytes_per_launch = algorithmic_bytes_per_scan(CHAINS_PER_GPU, DIM) * args.scans     # per GPU
    launch_ms = kernel_ms / args.steps
    achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
    n_local = pt.engine.n_local
    traffic, traffic_scans, traffic_src = read_ncu_traffic(CFG_NAME)
    line = {
        "metric": "pt_scans_per_s", "value": value, "unit": "scans/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": launch_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
        "ladder_scans_per_s": ladder_scans_per_s,
        "log_potential_evals_per_s": {"ref_equiv": evals / (kernel_ms * 1e-3), "unique_points": pts / (kernel_ms * 1e-3),
                                      "note": "ref_equiv = log_potential/logdensity[_and_gradient] calls the reference code "
                                              "path makes for the same trajectory; unique = density points the kernel evaluates"},
        "e2e": {"value": e2e_value, "unit": "scans/s",
                "h2d_bytes_per_step": n_chains * 8 + 96 + (DIM * 8 if CFG["explorer"] == "AutoMALA" else 0),
                "d2h_bytes_per_step": n_local * 120 + 2 * 128 * 8 + 12,
                "note": "wall clock around set_schedule + set_explorer + pgn_run_round (host buffers in, statistics out)"},
        "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_note": (f"DRAM bytes of one {traffic_scans}-scan launch, {traffic_src}; the replica "
                                                          "state stays in registers for the whole launch, so the traffic does not "
                                                          "grow with the number of scans" if traffic is not None else None),
                     "peak_source": peak_src,
                     "kernel": CFG["kernel"] + " (one persistent launch per step)",
                     "algorithmic_bytes_per_scan": algorithmic_bytes_per_scan(CHAINS_PER_GPU, DIM),
                     "note": "the working set is register-resident for the whole round: the path is bound by the dependency "
                             "latency along each replica's serial chain, not by HBM (SURVEY.md §8d)"},
        "clocks": clocks,
        "timed_region_s": t_region,
    }
    if CFG_NAME == "c5" and batch_steps > 0:
        flops_per_batch = 2 * 2.0 * CFG["n_data"] * CFG["dim"] * CHAINS_PER_GPU      # two GEMMs, 2 flops per fma
        fp64_peak = lib.measure_fp64_peak(local_rank)
        line["fp64_roofline"] = {
            "bound": "fp64 (tcgen05 has no FP64 path; DMMA m8n8k4 GEMM whose summation order equals the sequential-fma spec; "
                     "PGN_GEMM=simt selects the DFMA kernel)",
            "gemm": os.environ.get("PGN_GEMM", "dmma"),
            "achieved_tflops": flops_per_batch * batch_steps / (gemm_ms * 1e-3) / 1e12,
            "peak_tflops": fp64_peak, "peak_source": "measured in-run: register-resident DFMA loop on all SMs",
            "frac": flops_per_batch * batch_steps / (gemm_ms * 1e-3) / 1e12 / fp64_peak,
            "gemm_share_of_kernel_time": gemm_ms / kernel_ms, "batched_evaluations": batch_steps,
            "flops_per_batched_evaluation": flops_per_batch}
    if not args.no_cpu_baseline and args.gpus == 1 and CFG_NAME == "c5":
        line["cpu_baseline"] = cpu_baseline_c5(pg, args.cpu_threads)
    if not args.no_cpu_baseline and args.gpus == 1 and CFG_NAME != "c5":
        cpt, clib, n_threads, scans = cpu_reference_run(pg, args.scans, CHAINS_PER_GPU, threads=args.cpu_threads, budget_s=12.0,
                                                          clone_from=pt)
        r = one_step(cpt, pg, scans)
        line["cpu_baseline"] = {"value": scans / r.wall_s, "unit": "scans/s", "cores": n_threads, "kind": "port",
                                "sample": f"{scans} scans of the same {CHAINS_PER_GPU}-chain workload, continuing from the GPU arm's adapted state "
                                          "(oracle/: C++ restatement of the reference path, OpenMP over replicas)"}
    print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
