#!/usr/bin/env python
"""Attribute an ncu per-instruction SASS listing of nps_step_kernel to source files / functions.

    cuobjdump -xelf all libnps_b200.so ; nvdisasm -gi -c nps_capi.sm_100a.cubin > dis.txt
    ncu -i prof.ncu-rep --page source --csv > sass.csv
    python profiles/attribute_sass.py dis.txt sass.csv [n_warp_steps] [kernel symbol substring]

The ncu source page only carries metrics for the kernel's own .cu file; everything on the step path is
inlined from csrc/plant/*.h, so the join goes through nvdisasm's line table (same cubin, same order).
Output: executed warp-instructions and stall samples per header, per (header, enclosing source function).
"""
import collections
import csv
import re
import sys


def parse_dis(path, kernel="nps_step_kernel"):
    """-> list of (offset, opcode, file, line, outer_file, outer_line) for the kernel, in address order.
    nvdisasm -gi prints one '//## File' line per inline frame, innermost first, before each instruction group;
    (file, line) is the innermost frame, (outer_file, outer_line) the innermost frame inside csrc/plant that is not
    the hd.h helper header (i.e. the physics function the instruction belongs to)."""
    out = []
    on = False
    frames = []
    fresh = True
    rx_file = re.compile(r'//## File "([^"]+)", line (\d+)')
    rx_ins = re.compile(r'/\*([0-9a-f]{4,})\*/\s+(.*?);')
    for ln in open(path):
        if ln.startswith("//---") and ".text." in ln:
            on = kernel in ln
            continue
        if not on:
            continue
        lab = re.match(r'\s*(\$\S+):', ln)
        if lab:      # libdevice subroutine ($kernel$__internal_accurate_pow ...): it carries no line info of its own
            frames = [("<libdevice>/" + lab.group(1).split('$')[-1], 0)]
            fresh = True
            continue
        m = rx_file.search(ln)
        if m:
            if fresh:
                frames = []
                fresh = False
            frames.append((m.group(1), int(m.group(2))))
            continue
        m = rx_ins.search(ln)
        if m:
            fresh = True
            f, l = frames[0] if frames else ("?", 0)
            of, ol = f, l
            for ff, ll in frames:
                if "/csrc/" in ff and not ff.endswith("hd.h"):
                    of, ol = ff, ll
                    break
            out.append((int(m.group(1), 16), m.group(2).strip(), f, l, of, ol))
    return out


def function_table(path):
    """line -> enclosing function name for one header (restricted style: NPS_HD <ret> name(...) { at column 0)."""
    tbl = []
    rx = re.compile(r'^(?:template.*)?\s*(?:NPS_HD|static|inline|__device__|__global__).*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(')
    try:
        for i, ln in enumerate(open(path), 1):
            if ln[:1] not in (" ", "\t", "/", "}", "#", "\n"):
                m = rx.match(ln)
                if m:
                    tbl.append((i, m.group(1)))
    except OSError:
        pass
    return tbl


def enclosing(tbl, line):
    name = "?"
    for l, n in tbl:
        if l <= line:
            name = n
        else:
            break
    return name


def main():
    dis, sass = sys.argv[1], sys.argv[2]
    nws = float(sys.argv[3]) if len(sys.argv) > 3 else 16384.0
    ins = parse_dis(dis, sys.argv[4] if len(sys.argv) > 4 else "nps_step_kernel")
    rows = list(csv.reader(open(sass)))
    hdr = rows[1]
    i_exec, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    i_lsb = hdr.index("stall_long_sb") if "stall_long_sb" in hdr else None
    i_noi = hdr.index("stall_no_inst") if "stall_no_inst" in hdr else None
    body = [r for r in rows[2:] if len(r) > i_exec]
    if len(body) != len(ins):
        print(f"WARNING: {len(body)} ncu rows vs {len(ins)} disassembled instructions (different build?)")
    n = min(len(body), len(ins))
    per_file = collections.defaultdict(lambda: [0, 0, 0, 0])
    per_fn = collections.defaultdict(lambda: [0, 0, 0, 0])
    per_leaf = collections.defaultdict(lambda: [0, 0, 0, 0])
    tables = {}
    for k in range(n):
        off, op, f, l, of, ol = ins[k]
        r = body[k]
        ex, sm = int(r[i_exec]), int(r[i_samp])
        lsb = int(r[i_lsb]) if i_lsb is not None else 0
        noi = int(r[i_noi]) if i_noi is not None else 0
        # the innermost frame that is in OUR tree decides the subsystem; libdevice/math frames count as leaves
        short = lambda p: p.split("/")[-1]
        ours, oline = of, ol
        if ours not in tables:
            tables[ours] = function_table(ours)
        fn = enclosing(tables[ours], oline)
        for d, key in ((per_file, short(ours)), (per_fn, (short(ours), fn)), (per_leaf, short(f))):
            d[key][0] += ex; d[key][1] += sm; d[key][2] += lsb; d[key][3] += noi
    tot_ex = sum(v[0] for v in per_file.values()); tot_sm = sum(v[1] for v in per_file.values())
    print(f"total executed warp-instructions {tot_ex}  ({tot_ex / nws:.0f} per warp-substep), samples {tot_sm}")
    print("\n== by header of the outermost frame in csrc/ ==")
    for k, v in sorted(per_file.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:24s} inst {v[0] / nws:9.0f} ({100 * v[0] / tot_ex:5.1f}%)  samples {100 * v[1] / tot_sm:5.1f}%  long_sb {100 * v[2] / tot_sm:5.1f}%  no_inst {100 * v[3] / tot_sm:5.1f}%")
    print("\n== by function ==")
    for k, v in sorted(per_fn.items(), key=lambda kv: -kv[1][1])[:60]:
        print(f"{k[0]:20s} {k[1]:42s} inst {v[0] / nws:9.0f} ({100 * v[0] / tot_ex:5.1f}%)  samples {100 * v[1] / tot_sm:5.1f}%  long_sb {100 * v[2] / tot_sm:5.1f}%")
    print("\n== by innermost file (math library frames) ==")
    for k, v in sorted(per_leaf.items(), key=lambda kv: -kv[1][0])[:15]:
        print(f"{k:32s} inst {v[0] / nws:9.0f} ({100 * v[0] / tot_ex:5.1f}%)  samples {100 * v[1] / tot_sm:5.1f}%")


if __name__ == "__main__":
    main()
