#!/usr/bin/env python
"""Where a kernel's time goes, by source FUNCTION: joins the per-SASS-instruction warp-stall samples and shared-memory
wavefronts of an .ncu-rep (read on the CPU box) with the inline chains `nvdisasm -gi` prints for the same cubin.

    python tools/ncu_regions.py REPORT.ncu-rep [path/to/libldpc_b200.so]

Every SASS instruction is attributed to the OUTERMOST region function on its inline chain (REGIONS below: the channel
draw, the two decoder phases, the team reduction, the stage epilogue ...); what is left is the kernel body itself.
The library must be the build the report was captured from (same instruction offsets)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STALLS = ["stall_barrier", "stall_branch_resolving", "stall_dispatch", "stall_lg", "stall_long_sb", "stall_math", "stall_mio",
          "stall_no_inst", "stall_not_selected", "stall_selected", "stall_short_sb", "stall_wait", "stall_sleep", "stall_membar"]
REGIONS = ["team_draw", "team_draw_lanes", "warp_draw", "ov_producer_loop", "check_phase_tm", "var_phase_tm", "check_phase", "var_phase", "mc_var_phase",
           "team_reduce", "stage_finish_group", "tm_fill_lists", "team_generate_graph", "stage_v2c_resident", "reload_group",
           "e2_check_phase_async", "e2_var_phase_async", "e2_check_phase", "e2_var_phase", "e2_count_phase", "stage_reload_group", "mc_group"]


def functions_of(path, cache={}):
    """line -> name of the function whose body contains it (brace counting; good enough for this code base)"""
    if path in cache:
        return cache[path]
    out = {}
    try:
        src = open(path, errors="replace").read().split("\n")
    except OSError:
        cache[path] = out
        return out
    depth, cur, pending = 0, None, None
    for ln, text in enumerate(src, 1):
        code = re.sub(r"//.*", "", text)
        if depth <= 1:
            skip = ("if", "for", "while", "switch", "__launch_bounds__", "defined", "alignas", "asm", "e1_threads", "ov_threads")
            names = [nm for nm in re.findall(r"\b([A-Za-z_]\w*)\s*\(", code) if nm not in skip]
            if names and ";" not in code and not code.lstrip().startswith(("if", "for", "while", "switch", "return", "#", "else")) and \
                    pending is None and cur is None:
                pending = names[0]
        opens, closes = code.count("{"), code.count("}")
        if pending and opens and cur is None and depth <= 1 and "namespace" not in code:
            cur, cur_depth = pending, depth
            pending = None
        if cur:
            out[ln] = cur
        depth += opens - closes
        if cur and depth <= cur_depth:
            cur = None
        if ";" in code and depth <= 1 and not opens:
            pending = None
    cache[path] = out
    return out


def region_of(chain):
    """chain: [(file, line), ...] innermost first"""
    names = []
    for f, l in chain:
        names.append(functions_of(f).get(l, "?"))
    for nm in reversed(names):   # outermost first
        if nm in REGIONS:
            return nm
    return names[-1] if names else "?"


def disassemble(lib, kernel_mangled):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
    cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    for cb in cubins:
        txt = subprocess.run(["nvdisasm", "-gi", "-c", cb], capture_output=True, text=True).stdout
        key = ".text." + kernel_mangled + ":"
        if key not in txt:
            continue
        body = txt.split(key, 1)[1]
        body = body.split("\n//--------------------- ", 1)[0]
        insts, chain, fresh = [], [], True
        for line in body.split("\n"):
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
            if m:
                if fresh:
                    chain, fresh = [], False
                chain.append((m.group(1), int(m.group(2))))
                continue
            m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", line)
            if m:
                insts.append((int(m.group(1), 16), m.group(2).strip(), list(chain)))
                fresh = True
        return insts
    raise SystemExit("kernel not found in " + lib)


def main(rep, lib):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    kname = next(r[1] for r in rows if r and r[0] == "Kernel Name")
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    ix = {n: i for i, n in enumerate(hdr)}
    data = [r for r in rows[h + 1:] if len(r) == len(hdr)]
    base = int(data[0][ix["Address"]], 16)
    # mangled name: ask cu++filt in reverse -- simpler: match by demangling every .text symbol
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
    mangled = None
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        syms = subprocess.run(["cuobjdump", "-elf", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        for s in set(re.findall(r"\.text\.(_Z\w+)", syms)):
            dem = subprocess.run(["cu++filt", s], capture_output=True, text=True).stdout.strip()
            if dem.replace(" ", "").replace("ldpc::", "").replace("void", "") == kname.replace(" ", "").replace("ldpc::", "").replace("void", ""):
                mangled = s
    if not mangled:
        raise SystemExit("cannot find the mangled name of " + kname)
    insts = disassemble(lib, mangled)
    by_off = {o: (t, c) for o, t, c in insts}
    agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
    stalls = collections.defaultdict(collections.Counter)
    missing = 0
    for r in data:
        off = int(r[ix["Address"]], 16) - base
        if off not in by_off:
            missing += 1
            continue
        reg = region_of(by_off[off][1])
        a = agg[reg]
        a[0] += int(r[ix["# Samples"]] or 0)
        a[1] += int(r[ix["Instructions Executed"]] or 0)
        a[2] += int(r[ix["L1 Wavefronts Shared"]] or 0)
        a[3] += int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
        a[4] += int(r[ix["stall_barrier"]] or 0)
        for nm in STALLS:
            if nm in ix:
                stalls[reg][nm] += int(r[ix[nm]] or 0)
    tot = [sum(a[k] for a in agg.values()) for k in range(5)]
    print(f"kernel: {kname}")
    print(f"regions by warp-stall samples (total {tot[0]}; {missing} instructions unmatched)")
    print(f"  {'region':22} {'samples':>8} {'%':>6} {'warp-inst':>12} {'%':>6} {'smem wavefr.':>13} {'%':>6} {'ideal':>12} {'barrier smp':>11}")
    for reg, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"  {reg:22} {a[0]:8d} {100 * a[0] / max(tot[0], 1):6.1f} {a[1]:12d} {100 * a[1] / max(tot[1], 1):6.1f} "
              f"{a[2]:13d} {100 * a[2] / max(tot[2], 1):6.1f} {a[3]:12d} {a[4]:11d}")
    print("top stall reasons per region (share of the region's samples)")
    for reg, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if a[0] * 200 < tot[0]:
            continue
        top = stalls[reg].most_common(9)
        print(f"  {reg:22} " + "  ".join(f"{n[6:]} {100 * v / max(a[0], 1):.0f}%" for n, v in top))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "iib_project_ldpc_codes_b200", "libldpc_b200.so"))
