"""Loads oracle/liborc_pigeons.so behind pigeons.jl_b200's own ctypes marshalling
(prefix `orc_`).  Test infrastructure: only tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs use this."""
import hashlib
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liborc_pigeons.so")
ORACLE_STAMP = os.path.join(ORACLE_DIR, "liborc_pigeons.stamp")


def _cpu_fingerprint():
    """The oracle is built -march=native: a library built on another CPU model may use instructions this one lacks."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return hashlib.sha1(" ".join(sorted(line.split(":", 1)[1].split())).encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def build_oracle(force=False):
    import fcntl
    with open(os.path.join(ORACLE_DIR, ".build.lock"), "w") as lock:     # several ranks may arrive here together
        fcntl.flock(lock, fcntl.LOCK_EX)
        return _build_oracle_locked(force)


def _build_oracle_locked(force):
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("orc_engine.cpp", "orc_math.hpp", "Makefile")]
    stale = (not os.path.exists(ORACLE_SO)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in srcs)
    fp = _cpu_fingerprint()
    try:
        with open(ORACLE_STAMP) as f:
            other_cpu = f.read().strip() != fp
    except OSError:
        other_cpu = True
    if force or stale or other_cpu:
        subprocess.run(["make", "-C", ORACLE_DIR, "-B"], check=True, capture_output=True)
        with open(ORACLE_STAMP, "w") as f:
            f.write(fp)
    return ORACLE_SO


def load_oracle():
    import pigeons_jl_b200 as pg
    lib = pg.EngineLib(build_oracle(), prefix="orc_")
    import ctypes as C
    lib.lib.orc_set_threads.argtypes = [C.c_void_p, C.c_int]
    lib.lib.orc_get_threads.argtypes = [C.c_void_p]
    return lib
