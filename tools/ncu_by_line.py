#!/usr/bin/env python
"""Join an ncu SASS source-page CSV with nvdisasm line info: instructions executed,
thread-level efficiency and stall samples per CUDA source line.

    ncu -i prof.ncu-rep --page source --csv --print-source=sass > sass.csv
    python tools/ncu_by_line.py sass.csv open_pcc_metric_b200/libpccm.so pair_query_kernel [KInt]
    python tools/ncu_by_line.py sass.csv open_pcc_metric_b200/libpccm.so vx_epilogue_kernel ILb1ELb0    # one instantiation of several
"""
import collections
import csv
import re
import subprocess
import sys
import tempfile
import os


def line_map(so, func_substr, extra):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout
    out = {}
    cur_line = None
    active = False
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            active = func_substr in m.group(1) and all(e in m.group(1) for e in extra)
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out[int(m.group(1), 16)] = (cur_line, m.group(2).strip())
    return out


def main():
    sass_csv, so, func = sys.argv[1:4]
    extra = sys.argv[4:]
    lm = line_map(so, func, extra)
    rows = list(csv.reader(open(sass_csv)))
    # the CSV may hold several kernels: pick blocks whose name matches
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
    hdr = None
    use = False
    base = None
    total = [0.0, 0.0, 0.0]
    for r in rows:
        if r and r[0] == "Kernel Name":
            use = func in r[1] and all(e in r[1] for e in extra if not e.startswith("IL"))   # ("IL...": a mangled template argument list -- line map only)
            hdr = None
            base = None
            continue
        if not use:
            continue
        if hdr is None:
            hdr = {h: i for i, h in enumerate(r)}
            continue
        addr = int(r[hdr["Address"]], 16)
        if base is None:
            base = addr
        off = addr - base

        def f(k):
            try:
                return float(r[hdr[k]].replace(",", ""))
            except Exception:
                return 0.0
        key = lm.get(off, (None, "?"))[0]
        v = (f("Instructions Executed"), f("Thread Instructions Executed"), f("# Samples"))
        for i in range(3):
            agg[key][i] += v[i]
            total[i] += v[i]
    print(f"total warp-inst {total[0]:.3e}  thread-inst {total[1]:.3e}  avg active {total[1] / max(total[0], 1):.1f}  samples {total[2]:.0f}")
    print(f"{'file:line':34s} {'warp-inst%':>10s} {'active':>7s} {'samples%':>9s}")
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
        name = f"{key[0]}:{key[1]}" if key else "?"
        print(f"{name:34s} {100 * v[0] / total[0]:10.2f} {v[1] / max(v[0], 1):7.1f} {100 * v[2] / max(total[2], 1):9.2f}")


if __name__ == "__main__":
    main()
