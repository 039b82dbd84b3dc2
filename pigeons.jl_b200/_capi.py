"""ctypes binding of the C ABI declared in include/pigeons_b200.h.

This is the reference-side binding a maintainer would write (the Julia
equivalent — `ccall((:pgn_run_round, libpigeons_b200), Cint, ...)` — is shown in
INTEGRATION.md).  The product only ever loads `libpigeons_b200.so`; there is no
CPU fallback: if the CUDA library is missing or no GPU is usable, construction
fails loudly.

`EngineLib` is parametrised by (path, prefix) only so that the test-suite can
drive the CPU oracle (prefix ``orc_``) through the very same marshalling code;
nothing in this package references the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np

ABI_VERSION = 4

# return codes (include/pigeons_b200.h)
PGN_OK = 0
ERR_NAMES = {
    1: "PGN_ERR_INVALID", 2: "PGN_ERR_NO_DEVICE", 3: "PGN_ERR_CUDA", 4: "PGN_ERR_NAN_RATIO",
    5: "PGN_ERR_BAD_DENSITY", 6: "PGN_ERR_SLICE_MAX_ITER", 7: "PGN_ERR_STEP_UNDERFLOW",
    8: "PGN_ERR_NOT_POSITIVE", 9: "PGN_ERR_TIMEOUT",
}

TARGET_TOY_MVN, TARGET_FUNNEL, TARGET_GMM, TARGET_ISING, TARGET_LOGREG, TARGET_TEST_SWAPPER = 1, 2, 3, 4, 5, 6
TARGET_MIXED = 7
TARGET_UNID = 8
EXPLORER_NONE, EXPLORER_TOY, EXPLORER_SLICE, EXPLORER_AUTOMALA, EXPLORER_ISING_METROPOLIS, EXPLORER_MALA = 0, 1, 2, 3, 4, 5
EXPLORER_COMPOSE, EXPLORER_MIX = 6, 7
MAX_MIX = 4
PRECOND_IDENTITY, PRECOND_DIAGONAL, PRECOND_MIX_DIAGONAL = 0, 1, 2
RECORDERS_PER_REPLICA, RECORDERS_PER_CHAIN = 0, 1

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_u64p = C.POINTER(C.c_uint64)
_u8p = C.POINTER(C.c_uint8)


class pgn_config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("target_kind", C.c_int32), ("dim", C.c_int32), ("n_chains", C.c_int32),
        ("seed", C.c_int64), ("rank", C.c_int32), ("world_size", C.c_int32), ("device", C.c_int32),
        ("n_modes", C.c_int32), ("p", C.c_double * 8),
        ("means", _dp), ("log_weights", _dp), ("data_x", _dp), ("data_y", _dp),
        ("recorder_order", C.c_int32), ("n_chains_variational", C.c_int32),
    ]


class pgn_explorer_params(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("slice_w", C.c_double), ("slice_p", C.c_int32), ("slice_n_passes", C.c_int32),
        ("slice_max_iter", C.c_int32), ("n_refresh", C.c_int32), ("step_size", C.c_double),
        ("precond_kind", C.c_int32), ("mix_p0", C.c_double), ("mix_p01", C.c_double),
        ("std_devs", _dp), ("ising_n_steps", C.c_int32),
        ("n_steps", C.c_int32), ("step_kind", C.c_int32 * MAX_MIX),
        ("n_mix", C.c_int32), ("mix_n_refresh", C.c_int32 * MAX_MIX), ("mix_precond_kind", C.c_int32 * MAX_MIX),
        ("mix_step_size", C.c_double * MAX_MIX), ("mix_variant_p0", C.c_double * MAX_MIX),
        ("mix_variant_p01", C.c_double * MAX_MIX),
    ]


class pgn_round_out(C.Structure):
    _fields_ = [
        ("swap_n", _i64p), ("swap_mean", _dp), ("logsum_fwd", _dp), ("logsum_bwd", _dp),
        ("expl_acc_n", _i64p), ("expl_acc_mean", _dp), ("expl_n_steps", _i64p),
        ("am_n", _i64p), ("am_mean", _dp), ("rev_n", _i64p), ("rev_mean", _dp),
        ("n_tempered_restarts", C.c_int64), ("n_round_trips", C.c_int64),
        ("online_n", C.c_int64), ("online_mean", _dp), ("online_var", _dp),
        ("index_process", _i32p), ("swap_lr", _dp), ("swap_u", _dp), ("swap_accept", _u8p),
        ("target_trace", _dp),
        ("n_density_points", C.c_int64), ("n_ref_equiv_evals", C.c_int64), ("kernel_ms", C.c_double),
        ("gemm_ms", C.c_double), ("batch_steps", C.c_int64),
        ("n_launches", C.c_int64), ("active_columns", C.c_int64), ("gemm_columns", C.c_int64),
    ]


class pgn_replica_state(C.Structure):
    _fields_ = [("x", _dp), ("replica_index", _i32p), ("rng_counter", _u64p), ("round_trip_state", _i32p)]


class pgn_device_info_t(C.Structure):
    _fields_ = [("sm_major", C.c_int32), ("sm_minor", C.c_int32), ("n_sms", C.c_int32),
                ("global_mem_bytes", C.c_int64), ("max_resident_chains", C.c_int32), ("name", C.c_char * 128)]


# every symbol include/pigeons_b200.h declares (checked by tests/test_host_logic.py::test_capi_library_loads_and_exports_every_symbol)
DECLARED_SYMBOLS = [
    "pgn_abi_version", "pgn_create", "pgn_destroy", "pgn_free_string", "pgn_device_info", "pgn_local_range",
    "pgn_set_schedule", "pgn_set_explorer", "pgn_init_replicas", "pgn_get_state", "pgn_set_state",
    "pgn_run_round", "pgn_log_potential", "pgn_logdensity_and_gradient", "pgn_ipc_export", "pgn_ipc_attach",
    "pgn_peer_attach", "pgn_test_math", "pgn_measure_fp64_peak", "pgn_test_dmma", "pgn_hamiltonian_dynamics",
    "pgn_set_variational",
]


class EngineError(RuntimeError):
    def __init__(self, code: int, message: str):
        self.code = code
        super().__init__(f"{ERR_NAMES.get(code, code)}: {message}")


def default_library_path() -> str:
    here = os.path.dirname(os.path.abspath(__file__))
    return os.path.join(here, "csrc", "libpigeons_b200.so")


def _ptr(a: Optional[np.ndarray], ctype):
    if a is None:
        return C.cast(None, C.POINTER(ctype))
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(ctype))


class EngineLib:
    """A loaded engine library (`libpigeons_b200.so`)."""

    def __init__(self, path: Optional[str] = None, prefix: str = "pgn_"):
        path = path or default_library_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        self.path = path
        self.prefix = prefix
        self.lib = C.CDLL(path)
        for name in ("create", "destroy", "set_schedule", "set_explorer", "init_replicas", "get_state",
                     "set_state", "run_round", "log_potential", "logdensity_and_gradient", "local_range",
                     "test_math"):
            getattr(self.lib, prefix + name).restype = C.c_int
        getattr(self.lib, prefix + "free_string").restype = None
        getattr(self.lib, prefix + "free_string").argtypes = [C.c_void_p]

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def check(self, rc: int, err: C.c_char_p):
        if rc != PGN_OK:
            msg = err.value.decode() if err.value else ""
            if err.value is not None:
                self.fn("free_string")(C.cast(err, C.c_void_p))
            raise EngineError(rc, msg)

    def call(self, name, *args):
        err = C.c_char_p()
        rc = self.fn(name)(*args, C.byref(err))
        self.check(rc, err)

    def measure_fp64_peak(self, device: int = 0) -> float:
        v = C.c_double()
        f = self.fn("measure_fp64_peak")
        f.restype = C.c_int
        self.call("measure_fp64_peak", C.c_int32(device), C.byref(v))
        return v.value

    def test_dmma(self, a, b, c, device: int = 0) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.float64); b = np.ascontiguousarray(b, dtype=np.float64)
        c = np.ascontiguousarray(c, dtype=np.float64)
        n = a.shape[0]
        assert a.shape == (n, 8, 4) and b.shape == (n, 4, 8) and c.shape == (n, 8, 8)
        out = np.empty((n, 8, 8), dtype=np.float64)
        self.fn("test_dmma").restype = C.c_int
        self.call("test_dmma", C.c_int32(device), _ptr(a, C.c_double), _ptr(b, C.c_double), _ptr(c, C.c_double),
                  _ptr(out, C.c_double), C.c_int32(n))
        return out

    def test_math(self, op: int, values, seed: int = 1, replica_index: int = 1, device: int = 0) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.float64)
        n = v.size // 2 if op == 6 else v.size
        out = np.empty(n, dtype=np.float64)
        self.call("test_math", C.c_int32(device), C.c_int32(op), _ptr(v, C.c_double), _ptr(out, C.c_double),
                  C.c_int64(n), C.c_int64(seed), C.c_int32(replica_index))
        return out


@dataclass
class RoundResult:
    """Outputs of one `run_round` call (numpy views of pgn_round_out)."""
    n_scans: int
    first_chain: int
    swap_n: np.ndarray
    swap_mean: np.ndarray
    logsum_fwd: np.ndarray
    logsum_bwd: np.ndarray
    expl_acc_n: np.ndarray
    expl_acc_mean: np.ndarray
    expl_n_steps: np.ndarray
    am_n: np.ndarray
    am_mean: np.ndarray
    rev_n: np.ndarray
    rev_mean: np.ndarray
    n_tempered_restarts: int
    n_round_trips: int
    online_n: int
    online_mean: np.ndarray
    online_var: np.ndarray
    index_process: Optional[np.ndarray]
    swap_lr: Optional[np.ndarray]
    swap_u: Optional[np.ndarray]
    swap_accept: Optional[np.ndarray]
    target_trace: Optional[np.ndarray]
    n_density_points: int
    n_ref_equiv_evals: int
    kernel_ms: float
    wall_s: float = 0.0
    gemm_ms: float = 0.0
    batch_steps: int = 0
    n_launches: int = 0
    active_columns: int = 0
    gemm_columns: int = 0


class Engine:
    """One engine handle = one shard of the chain ladder on one GPU."""

    def __init__(self, lib: EngineLib, *, target_kind: int, dim: int, n_chains: int, seed: int,
                 p=(), means=None, log_weights=None, data_x=None, data_y=None, n_modes: int = 0,
                 rank: int = 0, world_size: int = 1, device: int = 0, recorder_order: int = RECORDERS_PER_REPLICA,
                 n_chains_variational: int = 0):
        self.lib = lib
        self.dim = int(dim)
        self.n_chains = int(n_chains)
        cfg = pgn_config()
        cfg.abi_version = ABI_VERSION
        cfg.target_kind = target_kind
        cfg.dim = dim
        cfg.n_chains = n_chains
        cfg.seed = seed
        cfg.rank, cfg.world_size, cfg.device = rank, world_size, device
        cfg.n_modes = n_modes
        cfg.recorder_order = recorder_order
        cfg.n_chains_variational = n_chains_variational
        self.n_chains_variational = n_chains_variational
        for i, v in enumerate(p):
            cfg.p[i] = float(v)
        self._keep = []
        for name, arr in (("means", means), ("log_weights", log_weights), ("data_x", data_x), ("data_y", data_y)):
            if arr is not None:
                a = np.ascontiguousarray(arr, dtype=np.float64)
                self._keep.append(a)
                setattr(cfg, name, _ptr(a, C.c_double))
        self._h = C.c_void_p()
        lib.call("create", C.byref(cfg), C.byref(self._h))
        fc, nl = C.c_int32(), C.c_int32()
        lib.fn("local_range")(self._h, C.byref(fc), C.byref(nl))
        self.first_chain, self.n_local = fc.value, nl.value

    def close(self):
        if self._h:
            self.lib.fn("destroy")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration ------------------------------------------------------
    def set_schedule(self, beta):
        b = np.ascontiguousarray(beta, dtype=np.float64)
        self.lib.call("set_schedule", self._h, _ptr(b, C.c_double), C.c_int32(b.size))

    def set_variational(self, mean, sd):
        """GaussianReference of the variational leg (GaussianReference.jl:4-54); None, None switches it off."""
        self.lib.fn("set_variational").restype = C.c_int
        if mean is None or sd is None:
            self.lib.call("set_variational", self._h, None, None)
            return
        m = np.ascontiguousarray(mean, dtype=np.float64)
        s = np.ascontiguousarray(sd, dtype=np.float64)
        if m.shape != (self.dim,) or s.shape != (self.dim,):
            raise ValueError(f"set_variational: mean and sd must have shape ({self.dim},)")
        self.lib.call("set_variational", self._h, _ptr(m, C.c_double), _ptr(s, C.c_double))

    def set_explorer(self, *, kind, slice_w=10.0, slice_p=20, slice_n_passes=3, slice_max_iter=1024,
                     n_refresh=0, step_size=1.0, precond_kind=PRECOND_IDENTITY, mix_p0=1.0 / 3.0,
                     mix_p01=1.0 / 3.0 + 1.0 / 3.0, std_devs=None, ising_n_steps=3, mix_variants=None, steps=None):
        ep = pgn_explorer_params()
        ep.kind = kind
        ep.slice_w, ep.slice_p, ep.slice_n_passes, ep.slice_max_iter = slice_w, slice_p, slice_n_passes, slice_max_iter
        ep.n_refresh, ep.step_size = n_refresh, step_size
        ep.precond_kind, ep.mix_p0, ep.mix_p01 = precond_kind, mix_p0, mix_p01
        sd = None
        if std_devs is not None:
            sd = np.ascontiguousarray(std_devs, dtype=np.float64)
            assert sd.size == self.dim
        ep.std_devs = _ptr(sd, C.c_double)
        ep.ising_n_steps = ising_n_steps
        ep.n_mix = 0
        if mix_variants:      # Mix of autoMALA kernels: [(n_refresh, step_size, precond_kind, p0, p01), ...]
            if not (2 <= len(mix_variants) <= MAX_MIX):
                raise ValueError(f"a device Mix has 2..{MAX_MIX} explorers")
            ep.n_mix = len(mix_variants)
            for v, (nr, ss, pk, q0, q01) in enumerate(mix_variants):
                ep.mix_n_refresh[v], ep.mix_step_size[v], ep.mix_precond_kind[v] = nr, ss, pk
                ep.mix_variant_p0[v], ep.mix_variant_p01[v] = q0, q01
        ep.n_steps = 0
        if steps:             # Compose / Mix program: [(kind, n_refresh, step_size, precond_kind, p0, p01), ...]
            if not (1 <= len(steps) <= MAX_MIX):
                raise ValueError(f"a device Compose / Mix has 1..{MAX_MIX} explorers")
            ep.n_steps = len(steps)
            for v, (k, nr, ss, pk, q0, q01) in enumerate(steps):
                ep.step_kind[v] = k
                ep.mix_n_refresh[v], ep.mix_step_size[v], ep.mix_precond_kind[v] = nr, ss, pk
                ep.mix_variant_p0[v], ep.mix_variant_p01[v] = q0, q01
        self.lib.call("set_explorer", self._h, C.byref(ep))

    def init_replicas(self):
        self.lib.call("init_replicas", self._h)

    # -- state ---------------------------------------------------------------
    def get_state(self):
        n, d = self.n_local, self.dim
        x = np.zeros((n, d), dtype=np.float64)
        ri = np.zeros(n, dtype=np.int32)
        ctr = np.zeros(n, dtype=np.uint64)
        rt = np.zeros(n, dtype=np.int32)
        st = pgn_replica_state(_ptr(x, C.c_double), _ptr(ri, C.c_int32), _ptr(ctr, C.c_uint64), _ptr(rt, C.c_int32))
        self.lib.call("get_state", self._h, C.byref(st))
        return {"x": x, "replica_index": ri, "rng_counter": ctr, "round_trip_state": rt}

    def set_state(self, x=None, replica_index=None, rng_counter=None, round_trip_state=None):
        # pgn_set_state copies n_local rows from every buffer it is given: a buffer of any other size would make the
        # library read past it (or silently assign the wrong replicas to chains), so the shapes are checked here
        def prep(a, dt, shape, name):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dt)
            if a.shape != shape:
                raise ValueError(f"set_state: {name} has shape {a.shape}, this shard holds {shape} "
                                 f"(chains {self.first_chain}..{self.first_chain + self.n_local - 1}, dim {self.dim})")
            return a
        n, d = self.n_local, self.dim
        x = prep(x, np.float64, (n, d), "x")
        ri = prep(replica_index, np.int32, (n,), "replica_index")
        ctr = prep(rng_counter, np.uint64, (n,), "rng_counter")
        rt = prep(round_trip_state, np.int32, (n,), "round_trip_state")
        if ri is not None and (ri.min(initial=1) < 1 or ri.max(initial=1) > self.n_chains):
            raise ValueError("set_state: replica_index entries must lie in 1..n_chains")
        st = pgn_replica_state(_ptr(x, C.c_double), _ptr(ri, C.c_int32), _ptr(ctr, C.c_uint64), _ptr(rt, C.c_int32))
        self.lib.call("set_state", self._h, C.byref(st))

    # -- the hot path ----------------------------------------------------------
    def run_round(self, n_scans: int, *, log_index_process=False, log_swaps=False, log_target_trace=False) -> RoundResult:
        import time
        n, d = self.n_local, self.dim
        f64 = lambda *s: np.zeros(s, dtype=np.float64)   # noqa: E731
        i64 = lambda *s: np.zeros(s, dtype=np.int64)     # noqa: E731
        res = RoundResult(
            n_scans=n_scans, first_chain=self.first_chain,
            swap_n=i64(n), swap_mean=f64(n), logsum_fwd=f64(n), logsum_bwd=f64(n),
            expl_acc_n=i64(n), expl_acc_mean=f64(n), expl_n_steps=i64(n), am_n=i64(n), am_mean=f64(n),
            rev_n=i64(n), rev_mean=f64(n), n_tempered_restarts=0, n_round_trips=0, online_n=0,
            online_mean=f64(d), online_var=f64(d),
            index_process=np.zeros((n_scans, n), dtype=np.int32) if log_index_process else None,
            swap_lr=f64(n_scans, n) if log_swaps else None,
            swap_u=f64(n_scans, n) if log_swaps else None,
            swap_accept=np.zeros((n_scans, n), dtype=np.uint8) if log_swaps else None,
            target_trace=(f64(n_scans, 2, d) if 0 < self.n_chains_variational < self.n_chains else f64(n_scans, d))
            if log_target_trace else None,
            n_density_points=0, n_ref_equiv_evals=0, kernel_ms=0.0)
        out = pgn_round_out()
        out.swap_n, out.swap_mean = _ptr(res.swap_n, C.c_int64), _ptr(res.swap_mean, C.c_double)
        out.logsum_fwd, out.logsum_bwd = _ptr(res.logsum_fwd, C.c_double), _ptr(res.logsum_bwd, C.c_double)
        out.expl_acc_n, out.expl_acc_mean = _ptr(res.expl_acc_n, C.c_int64), _ptr(res.expl_acc_mean, C.c_double)
        out.expl_n_steps = _ptr(res.expl_n_steps, C.c_int64)
        out.am_n, out.am_mean = _ptr(res.am_n, C.c_int64), _ptr(res.am_mean, C.c_double)
        out.rev_n, out.rev_mean = _ptr(res.rev_n, C.c_int64), _ptr(res.rev_mean, C.c_double)
        out.online_mean, out.online_var = _ptr(res.online_mean, C.c_double), _ptr(res.online_var, C.c_double)
        out.index_process = _ptr(res.index_process, C.c_int32)
        out.swap_lr, out.swap_u = _ptr(res.swap_lr, C.c_double), _ptr(res.swap_u, C.c_double)
        out.swap_accept = _ptr(res.swap_accept, C.c_uint8)
        out.target_trace = _ptr(res.target_trace, C.c_double)
        t0 = time.perf_counter()
        self.lib.call("run_round", self._h, C.c_int64(n_scans), C.byref(out))
        res.wall_s = time.perf_counter() - t0
        res.n_tempered_restarts, res.n_round_trips = out.n_tempered_restarts, out.n_round_trips
        res.online_n = out.online_n
        res.n_density_points, res.n_ref_equiv_evals = out.n_density_points, out.n_ref_equiv_evals
        res.kernel_ms = out.kernel_ms
        res.gemm_ms, res.batch_steps = out.gemm_ms, out.batch_steps
        res.n_launches, res.active_columns, res.gemm_columns = out.n_launches, out.active_columns, out.gemm_columns
        return res

    # -- parity entry points -------------------------------------------------------
    def log_potential(self, x, beta) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, self.dim)
        b = np.ascontiguousarray(np.broadcast_to(beta, (x.shape[0],)), dtype=np.float64)
        out = np.empty(x.shape[0], dtype=np.float64)
        self.lib.call("log_potential", self._h, _ptr(x, C.c_double), C.c_int32(x.shape[0]),
                      _ptr(b, C.c_double), _ptr(out, C.c_double))
        return out

    def logdensity_and_gradient(self, x, beta):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, self.dim)
        b = np.ascontiguousarray(np.broadcast_to(beta, (x.shape[0],)), dtype=np.float64)
        ld = np.empty(x.shape[0], dtype=np.float64)
        g = np.empty_like(x)
        self.lib.call("logdensity_and_gradient", self._h, _ptr(x, C.c_double), C.c_int32(x.shape[0]),
                      _ptr(b, C.c_double), _ptr(ld, C.c_double), _ptr(g, C.c_double))
        return ld, g

    def hamiltonian_dynamics(self, x, p, beta, step_size: float, n_steps: int, diag_precond=None):
        """n_steps x leap_frog! (hamiltonian_dynamics.jl:39-93) by the engine's own integrator; diag_precond None = identity."""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, self.dim)
        p = np.ascontiguousarray(p, dtype=np.float64).reshape(-1, self.dim)
        b = np.ascontiguousarray(np.broadcast_to(beta, (x.shape[0],)), dtype=np.float64)
        xo, po = np.empty_like(x), np.empty_like(p)
        self.lib.fn("hamiltonian_dynamics").restype = C.c_int
        pre = None if diag_precond is None else np.ascontiguousarray(diag_precond, dtype=np.float64)
        if pre is not None and pre.shape != (self.dim,):
            raise ValueError(f"diag_precond must have shape ({self.dim},)")
        self.lib.call("hamiltonian_dynamics", self._h, _ptr(x, C.c_double), _ptr(p, C.c_double), C.c_int32(x.shape[0]),
                      _ptr(b, C.c_double), _ptr(pre, C.c_double), C.c_double(step_size), C.c_int32(n_steps), _ptr(xo, C.c_double), _ptr(po, C.c_double))
        return xo, po

    # -- multi-GPU ---------------------------------------------------------------------
    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self.lib.call("ipc_export", self._h, buf)
        return buf.raw

    def ipc_attach(self, side: int, handle: bytes):
        buf = C.create_string_buffer(handle, 64)
        self.lib.call("ipc_attach", self._h, C.c_int32(side), buf)

    def peer_attach(self, side: int, other: "Engine"):
        self.lib.call("peer_attach", self._h, C.c_int32(side), other._h)
