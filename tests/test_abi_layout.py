"""The three descriptions of the C ABI — include/pigeons_b200.h (as compiled by gcc: tests/abi_layout.c prints
sizeof/offsetof of every struct), the ctypes mirror (pigeons.jl_b200/_capi.py) and the Julia mirror
(julia/PigeonsB200.jl) — must agree field for field.  Julia cannot run in the build container, so its struct
blocks are parsed and laid out with the C rules (natural alignment), which is what `ccall` does for `isbits` structs."""
import ctypes as C
import json
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def layout():
    exe = os.path.join(ROOT, "tests", "abi_layout.bin")
    src = os.path.join(ROOT, "tests", "abi_layout.c")
    hdr = os.path.join(ROOT, "include", "pigeons_b200.h")
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["gcc", "-std=c11", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", exe, src], check=True)
    return json.loads(subprocess.run([exe], check=True, capture_output=True, text=True).stdout)


def test_probe_lists_every_field_of_the_header(layout):
    """abi_layout.c names its fields by hand: make sure none of the header's is missing."""
    hdr = open(os.path.join(ROOT, "include", "pigeons_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    for name, info in layout["structs"].items():
        body = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s;" % (name, name), hdr, flags=re.S).group(1)
        decl = []
        for stmt in body.split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            for part in stmt.split(","):
                m = re.search(r"([A-Za-z_][A-Za-z_0-9]*)\s*(\[[^\]]*\])?\s*$", part.strip())
                decl.append(m.group(1))
        assert decl == [f[0] for f in info["fields"]], name


def test_ctypes_mirror_matches_the_header(layout):
    from pigeons_jl_b200 import _capi
    assert layout["abi_version"] == _capi.ABI_VERSION
    for name, info in layout["structs"].items():
        cls = getattr(_capi, name)
        assert C.sizeof(cls) == info["size"], name
        assert [f[0] for f in cls._fields_] == [f[0] for f in info["fields"]], name
        for fname, off, size in info["fields"]:
            fld = getattr(cls, fname)
            assert (fld.offset, fld.size) == (off, size), f"{name}.{fname}"


JL_TYPES = {"Int32": (4, 4), "Int64": (8, 8), "UInt64": (8, 8), "Float64": (8, 8), "UInt8": (1, 1)}
JL_STRUCTS = {"PgnConfig": "pgn_config", "PgnExplorerParams": "pgn_explorer_params", "PgnRoundOut": "pgn_round_out",
              "PgnReplicaState": "pgn_replica_state", "PgnDeviceInfo": "pgn_device_info_t"}


def jl_size_align(t):
    t = t.strip()
    if t.startswith("Ptr{"):
        return 8, 8
    m = re.match(r"NTuple\{(\d+),\s*(\w+)\}", t)
    if m:
        s, a = JL_TYPES[m.group(2)]
        return int(m.group(1)) * s, a
    return JL_TYPES[t]


def test_julia_mirror_matches_the_header(layout):
    jl = open(os.path.join(ROOT, "julia", "PigeonsB200.jl")).read()
    assert int(re.search(r"const PGN_ABI_VERSION = (\d+)", jl).group(1)) == layout["abi_version"]
    hdr = open(os.path.join(ROOT, "include", "pigeons_b200.h")).read()
    for const, val in re.findall(r"^const (PGN_[A-Z_0-9]+) = (\d+)\s*$", jl, flags=re.M):      # every numeric constant equals the header's
        m = re.search(r"#define %s (\d+)" % const, hdr)
        assert m and int(m.group(1)) == int(val), const
    for jname, cname in JL_STRUCTS.items():
        body = re.search(r"^struct %s\n(.*?)^end" % jname, jl, flags=re.S | re.M).group(1)
        fields = [ln.strip().split("::") for ln in body.strip().splitlines() if "::" in ln]
        off, max_align, got = 0, 1, []
        for fname, ftype in fields:
            size, align = jl_size_align(ftype)
            off = (off + align - 1) // align * align
            got.append([fname, off, size])
            off += size
            max_align = max(max_align, align)
        total = (off + max_align - 1) // max_align * max_align
        info = layout["structs"][cname]
        assert got == info["fields"], f"{jname} vs {cname}"
        assert total == info["size"], jname
