"""SASS evidence for the hot kernels: opcode histogram of each kernel and of its straight-line (no-boundary) path,
spill instructions, and the path's listing.  Reads the objects of the last build (BUILD dir of xlb_b200/csrc/Makefile).

    python scripts/sass_evidence.py [/tmp/xlb_b200_build] > profiles/r2_sass_evidence.txt
"""

import collections
import re
import subprocess
import sys

BUILD = sys.argv[1] if len(sys.argv) > 1 else "/tmp/xlb_b200_build"
KERNELS = [  # (title, object, mangled-name fragment)
    ("D3Q19 BGK FP32FP32, scalar tile kernel (TMA-fed, persistent, one cell per thread): the headline kernel", "step_inst_d3q19_bgk.o", "step_tile1_kernelINS_7LatticeINS_9D3Q19BaseEEELi0EffLi1EEEv"),
    ("D3Q19 BGK FP32FP32, direct-load kernel, one cell per thread (slab faces, shapes that cannot be tiled)", "step_inst_d3q19_bgk.o", "step_kernelINS_7LatticeINS_9D3Q19BaseEEELi0EffLi1ELi0EEEv"),
    ("D3Q19 BGK FP32FP16, half2-state pair path (direct loads)", "step_inst_d3q19_bgk.o", "step_kernelINS_7LatticeINS_9D3Q19BaseEEELi0Ef6__halfLi2ELi2EEEv"),
    ("D3Q19 BGK FP32FP16, tile kernel (TMA-fed, persistent), 1024-cell tiles, 1 CTA/SM", "step_inst_d3q19_bgk.o", "step_tile_kernelINS_7LatticeINS_9D3Q19BaseEEELi1024ELi1EEEv"),
    ("D3Q27 KBC FP32FP32, register-lean formulation (default)", "step_inst_kbc_lean.o", "step_kernelINS_7LatticeINS_9D3Q27BaseEEELi9EffLi1ELi0EEEv"),
]


def sass(obj, frag):
    out = subprocess.run(["cuobjdump", "-sass", f"{BUILD}/{obj}"], capture_output=True, text=True).stdout
    lines, on = [], False
    for line in out.splitlines():
        if "Function :" in line:
            on = frag in line
            continue
        if on and re.search(r"/\*[0-9a-f]{4,5}\*/", line):
            lines.append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).rstrip())
    return lines


def opcode(line):
    m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
    return m.group(1) if m else "?"


def hist(lines, top=14):
    c = collections.Counter(opcode(l).split(".")[0] for l in lines)
    return "  ".join(f"{k} {v}" for k, v in c.most_common(top)), c


def regs(obj, frag):
    out = subprocess.run(["cuobjdump", "-res-usage", f"{BUILD}/{obj}"], capture_output=True, text=True).stdout.splitlines()
    for i, line in enumerate(out):
        if frag in line and i + 1 < len(out):
            return out[i + 1].strip()
    return "?"


def straight_path(lines):
    """The no-boundary path of the interior-plane variant: among the stretches between UNCONDITIONAL control transfers (EXIT / RET / BRA
    without a predicate; predicated branches and the CALL of the division slow path stay inside) that hold a full set of population
    loads and stores, the one with the fewest local-memory instructions, then the shortest."""
    segs, start = [], 0
    for i, l in enumerate(lines):
        body = l.split("*/", 1)[-1].strip()
        if re.match(r"(EXIT|RET|BRA)\b", body):
            segs.append(lines[start : i + 1])
            start = i + 1
    good = [g for g in segs if sum("STG" in x for x in g) >= 9 and sum(("LDG" in x or "LDS" in x) for x in g) >= 9]
    if not good:
        # persistent kernels: the loads (LDS out of the stage) and the stores of the loop body can sit in different stretches; take the
        # stretch with the most population loads and the cheapest stretch with a full set of stores, in program order
        loads = max(segs, key=lambda g: sum(("LDS" in x or "LDG" in x) for x in g), default=[])
        stores = [g for g in segs if sum("STG" in x for x in g) >= 9]
        if not stores or loads is None:
            return []
        st = min(stores, key=lambda g: (sum(("STL" in x or "LDL" in x) for x in g), len(g)))
        return loads + ([] if st is loads else st)
    return min(good, key=lambda g: (sum(("STL" in x or "LDL" in x) for x in g), len(g)))


for title, obj, frag in KERNELS:
    lines = sass(obj, frag)
    if not lines:
        print(f"== {title}: not found in {obj}\n")
        continue
    h, c = hist(lines)
    print(f"== {title}\n   {frag}\n   {regs(obj, frag)}")
    print(f"   whole kernel ({len(lines)} instructions, all x-plane variants and boundary paths): {h}")
    print(f"   local memory in the whole kernel: STL {c.get('STL', 0)}  LDL {c.get('LDL', 0)}")
    path = straight_path(lines)
    hp, cp = hist(path)
    print(f"   straight-line path ({len(path)} instructions): {hp}")
    print(f"   local memory on the path: STL {cp.get('STL', 0)}  LDL {cp.get('LDL', 0)};  FFMA {cp.get('FFMA', 0)}  FFMA2 {cp.get('FFMA2', 0)}  FMUL2 {cp.get('FMUL2', 0)}  FADD2 {cp.get('FADD2', 0)}  "
          f"LDG {cp.get('LDG', 0)}  LDS {cp.get('LDS', 0)}  STG {cp.get('STG', 0)}  UBLKCP {c.get('UBLKCP', 0)} (whole kernel)  SYNCS {c.get('SYNCS', 0)} (whole kernel)")
    print("   --- path listing ---")
    for l in path:
        print("   " + re.sub(r"\s+", " ", l.split("*/", 1)[-1]).strip())
    print()
