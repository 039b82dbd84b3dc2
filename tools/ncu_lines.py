#!/usr/bin/env python
"""Per-source-line summary of an ncu report (no GPU needed).

ncu's `--page source --csv` lists SASS instructions with their executed counts and stall samples but
not the CUDA line they came from; `nvdisasm -gi` knows the lines (the library is built with
-lineinfo).  This joins the two by instruction order and prints, per source line, the executed
warp-instructions and stall samples — by innermost line and by the line of a chosen outer function
(the call site inside it), so a hot inlined helper is attributed to its callers too.

usage: ncu_lines.py REPORT.ncu-rep|SOURCE_PAGE.csv MANGLED_KERNEL_SUBSTRING [--so LIB] [--top N] [--outer LO:HI]
"""
import argparse
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sh(cmd, **kw):
    return subprocess.run(cmd, check=True, capture_output=True, text=True, **kw).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel", help="substring of the mangled kernel name, e.g. VecChainILi2ELi1ELi3E")
    ap.add_argument("--so", default=os.path.join(ROOT, "pigeons.jl_b200", "csrc", "libpigeons_b200.so"))
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--outer", default=None, help="LO:HI line range of an outer function to attribute call sites to")
    a = ap.parse_args()

    # The library is linked from several translation units (csrc/Makefile), several of them built from the same source
    # with different -D flags: extract the cubin of every object file into its own directory and find the kernel.
    tmp = tempfile.mkdtemp()
    build = os.path.join(os.path.dirname(os.path.abspath(a.so)), "build")
    objs = sorted(os.path.join(build, f) for f in os.listdir(build) if f.endswith(".o")) if os.path.isdir(build) else [os.path.abspath(a.so)]
    idx = cubin = name = None
    for n, obj in enumerate(objs):
        sub = os.path.join(tmp, str(n))
        os.makedirs(sub)
        subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=sub, capture_output=True, text=True)
        for f in sorted(os.listdir(sub)):
            if not f.endswith(".cubin"):
                continue
            cand = os.path.join(sub, f)
            syms = subprocess.run(["readelf", "-sW", cand], capture_output=True, text=True).stdout
            for ln in syms.splitlines():
                if " FUNC " in ln and " GLOBAL " in ln and a.kernel in ln:
                    idx, cubin, name = int(ln.split(":")[0]), cand, ln.split()[-1]
                    break
            if idx is not None:
                break
        if idx is not None:
            break
    if idx is None:
        sys.exit("kernel not found in the cubins of " + a.so)
    sass = subprocess.run(["nvdisasm", "-gi", "-fun", str(idx), cubin], capture_output=True, text=True).stdout
    start = sass.index(".text." + name + ":")
    insts = []          # (addr, text, [lines innermost..outermost])
    frames = []
    re_file = re.compile(r'//## File "([^"]+)", line (\d+)')
    re_inst = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);")
    for ln in sass[start:].splitlines():
        m = re_file.search(ln)
        if m:
            frames.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re_inst.search(ln)
        if m:
            if frames:
                cur = frames
            insts.append((int(m.group(1), 16), m.group(2).strip(), cur if insts or frames else []))
            frames = []
        elif ln.startswith("//---") and insts:
            break
    # REPORT is an .ncu-rep, or the CSV already exported from one with `ncu -i X.ncu-rep --page source --csv`
    rep = open(a.report).read() if a.report.endswith(".csv") else sh(["ncu", "-i", a.report, "--page", "source", "--csv"])
    rows = list(csv.reader(io.StringIO(rep)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    col = {n: i for i, n in enumerate(rows[h])}
    data = rows[h + 1:]
    if len(data) != len(insts):
        print(f"warning: {len(data)} profiled instructions vs {len(insts)} disassembled", file=sys.stderr)
    n = min(len(data), len(insts))
    tot_ex = sum(float(data[i][col["Instructions Executed"]]) for i in range(n))
    tot_st = sum(float(data[i][col["# Samples"]]) for i in range(n))
    inner, outer = {}, {}
    lo, hi = (int(v) for v in a.outer.split(":")) if a.outer else (0, 0)
    for i in range(n):
        ex = float(data[i][col["Instructions Executed"]])
        st = float(data[i][col["# Samples"]])
        fr = insts[i][2]
        k = fr[0] if fr else ("?", 0)
        e = inner.setdefault(k, [0.0, 0.0])
        e[0] += ex; e[1] += st
        if a.outer:
            site = next((f for f in fr if f[0] == "pgn_kernels.cuh" and lo <= f[1] <= hi), ("(outside)", 0))
            e = outer.setdefault(site, [0.0, 0.0])
            e[0] += ex; e[1] += st
    src = {}
    for f in set(k[0] for k in inner):
        p = os.path.join(ROOT, "pigeons.jl_b200", "csrc", f)
        if os.path.exists(p):
            src[f] = open(p).read().splitlines()

    def show(tbl, title):
        print(f"== {title}: executed warp-instructions {tot_ex:.3g}, stall samples {tot_st:.0f}")
        for k, (ex, st) in sorted(tbl.items(), key=lambda kv: -kv[1][1])[:a.top]:
            text = src.get(k[0], [""] * (k[1] + 1))[k[1] - 1].strip()[:90] if k[1] > 0 and k[0] in src else ""
            print(f"{k[0]}:{k[1]:<5d} inst {100 * ex / tot_ex:5.1f}%  stall {100 * st / tot_st:5.1f}%  | {text}")

    show(inner, "by innermost line")
    if a.outer:
        show(outer, f"by call site inside lines {a.outer}")


if __name__ == "__main__":
    main()
