"""Establishes the accumulation order of mma.sync.m8n8k4.f64 on this GPU by comparing the
hardware result with exactly-rounded candidate orders (rational arithmetic).

    python tools/probe_dmma_order.py            # prints the match rate of every candidate
"""
import os
import sys
from fractions import Fraction

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np                 # noqa: E402
import pigeons_jl_b200 as pg       # noqa: E402


def fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def candidates(a, b, c):
    """a: [4] (row of A), b: [4] (column of B), c scalar."""
    out = {}
    acc = c
    for k in range(4):
        acc = fma(a[k], b[k], acc)
    out["fma chain k=0..3 from c"] = acc
    acc = c
    for k in reversed(range(4)):
        acc = fma(a[k], b[k], acc)
    out["fma chain k=3..0 from c"] = acc
    out["single rounding of exact sum"] = float(sum(Fraction(a[k]) * Fraction(b[k]) for k in range(4)) + Fraction(c))
    acc = 0.0
    for k in range(4):
        acc = fma(a[k], b[k], acc)
    out["fma chain from 0 then + c"] = acc + c
    p01 = fma(a[1], b[1], a[0] * b[0])
    p23 = fma(a[3], b[3], a[2] * b[2])
    out["pairwise (01)+(23)+c"] = (p01 + p23) + c
    prods = [a[k] * b[k] for k in range(4)]
    out["rounded products summed in order from c"] = (((c + prods[0]) + prods[1]) + prods[2]) + prods[3]
    return out


def main():
    rng = np.random.default_rng(0)
    n = 64
    lib = pg.EngineLib()
    scale = lambda shape: rng.normal(0, 1, shape) * np.exp(rng.uniform(-8, 8, shape))   # noqa: E731
    a, b, c = scale((n, 8, 4)), scale((n, 4, 8)), scale((n, 8, 8))
    d = lib.test_dmma(a, b, c)
    counts, total = {}, 0
    for t in range(n):
        for i in range(8):
            for j in range(8):
                cand = candidates(a[t, i, :], b[t, :, j], c[t, i, j])
                total += 1
                for k, v in cand.items():
                    counts[k] = counts.get(k, 0) + (1 if v == d[t, i, j] else 0)
    for k, v in sorted(counts.items(), key=lambda kv: -kv[1]):
        print(f"{v:6d}/{total}  {k}")
    ok = counts["fma chain k=0..3 from c"] == total
    print("DMMA == sequential fma chain (k ascending):", ok)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
