"""Loads oracle/liborc_pigeons.so behind pigeons.jl_b200's own ctypes marshalling
(prefix `orc_`).  Test infrastructure: only tests/, smoke() and bench.py's
cpu_baseline / --impl reference legs use this."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liborc_pigeons.so")


def build_oracle(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("orc_engine.cpp", "orc_math.hpp")]
    stale = (not os.path.exists(ORACLE_SO)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return ORACLE_SO


def load_oracle():
    import pigeons_jl_b200 as pg
    lib = pg.EngineLib(build_oracle(), prefix="orc_")
    import ctypes as C
    lib.lib.orc_set_threads.argtypes = [C.c_void_p, C.c_int]
    lib.lib.orc_get_threads.argtypes = [C.c_void_p]
    return lib
