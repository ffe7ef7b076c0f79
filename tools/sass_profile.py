"""Development helper: joins an `ncu --page source --csv` SASS dump of k_frame_energy with `nvdisasm -gi` line info of the
same cubin and prints instruction / stall-sample shares per kernel phase and per inlined device function.

usage: sass_profile.py <ncu_source.csv> <nvdisasm_gi_listing_of_the_kernel.txt> [frames warps]
"""
import collections
import csv
import re
import sys

src_csv, dis = sys.argv[1], sys.argv[2]
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 1184
warps = int(sys.argv[4]) if len(sys.argv) > 4 else 16

# ---- line info per instruction offset
info = {}
chain = []
fresh = True
pat_i = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);")
pat_f = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
for line in open(dis):
    m = pat_f.search(line)
    if m:
        if fresh:
            chain = []
            fresh = False
        chain.append((m.group(1).split("/")[-1], int(m.group(2)), m.group(3).split("/")[-1] if m.group(3) else None, int(m.group(4)) if m.group(4) else None))
        continue
    m = pat_i.match(line)
    if m:
        fresh = True
        info[int(m.group(1), 16)] = list(chain)


def classify(ch):
    """-> (outer frame_kernels.cu line, innermost repo file:line, path of repo functions)"""
    outer = None
    inner = None
    for f, l, pf, pl in ch:
        if inner is None and not f.endswith(".hpp") and not f.endswith(".h"):
            inner = (f, l)
        if pf == "frame_kernels.cu":
            outer = pl
        if f == "frame_kernels.cu" and pf is None:
            outer = l
    if outer is None and ch:
        f, l, pf, pl = ch[-1]
        outer = l if f == "frame_kernels.cu" else None
    return outer, inner


def _phases():
    """line ranges of the kernel's phases, from the marker comments of the current frame_kernels.cu"""
    import os
    src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mythos_b200", "csrc", "frame_kernels.cu")).read().split("\n")
    marks = [("__global__ void __launch_bounds__(kFB, 1) k_frame_energy", "prologue"), ("// persistent CTA: frames", "stage"),
             ("phase B: bonded pairs", "bonded"), ("-- unbonded pairs", "setup"), ("all-pairs mode: shared-memory cell list", "cells"),
             ("-- scheduler loop", "loop-head"), ("phase 3a:", "hb"), ("phase 3b:", "cross"), ("phase 3c:", "coax"),
             ("phase 2:", "phase2-exc"), ("phase 1:", "phase1-debye"), ("if (flush) break;", "producer"),
             ("#ifdef MB_FRAME_PROFILE\n  if (threadIdx.x == 0 && blockIdx.x == 0)", "flush"), ("fused observables", "observables"),
             ("static bool pick_layout", "end")]
    found = []
    for text, name in marks:
        first = text.split("\n")[0]
        for n, line in enumerate(src, 1):
            if first in line and (not found or n > found[-1][0]):
                found.append((n, name))
                break
    return [(found[k][0], found[k + 1][0] - 1, found[k][1]) for k in range(len(found) - 1)]


PHASES = _phases()


def phase_of(line):
    if line is None:
        return "?"
    for lo, hi, name in PHASES:
        if lo <= line <= hi:
            return name
    if line < 321:
        return "helpers(q_push..)"
    return "?"


rows = list(csv.reader(open(src_csv)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0


base = int(data[0][ix["Address"]], 16)
tot_i = sum(f(r, "Instructions Executed") for r in data)
tot_s = sum(f(r, "# Samples") for r in data)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ph_i, ph_s = collections.Counter(), collections.Counter()
ph_st = collections.defaultdict(collections.Counter)
fn_i, fn_s = collections.Counter(), collections.Counter()
ph_op = collections.defaultdict(collections.Counter)
last_outer = None
for r in data:
    off = int(r[ix["Address"]], 16) - base
    outer, inner = classify(info.get(off, []))
    if outer is None:
        outer = last_outer
    last_outer = outer
    ph = phase_of(outer)
    ie, sm = f(r, "Instructions Executed"), f(r, "# Samples")
    ph_i[ph] += ie
    ph_s[ph] += sm
    for c in stall_cols:
        ph_st[ph][c] += f(r, c)
    key = (ph, inner[0] + ":" + str(inner[1]) if inner else "?")
    fn_i[key] += ie
    fn_s[key] += sm
    s = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip())
    ph_op[ph][s.split()[0].split(".")[0]] += ie

print(f"dynamic warp instructions {tot_i:.3e} ({tot_i / frames / warps:.0f} per warp per frame), samples {tot_s:.0f}")
for ph, c in ph_i.most_common():
    st = ph_st[ph]
    top = ", ".join(f"{k[6:]} {v / max(ph_s[ph], 1) * 100:.0f}%" for k, v in st.most_common(5))
    print(f"{ph:18s} inst {c / tot_i * 100:5.1f}%  samples {ph_s[ph] / tot_s * 100:5.1f}%  inst/warp/frame {c / frames / warps:8.0f}   stalls: {top}")
    print("                   ops: " + ", ".join(f"{o} {v / max(c, 1) * 100:.0f}%" for o, v in ph_op[ph].most_common(10)))
print()
print("top source lines by samples:")
for (ph, k), s in fn_s.most_common(60):
    print(f"  {ph:14s} {k:34s} samples {s / tot_s * 100:5.2f}%  inst {fn_i[(ph, k)] / tot_i * 100:5.2f}%")
