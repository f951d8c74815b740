#!/usr/bin/env python
"""Build A/B variants of libpccm.so (extra -D flags) into build/ and, on a GPU box, print the ncu launch list of
one bench step for each:   python tools/variants.py build name1:-DX=1,-DY=2 name2:...   |   python tools/variants.py run name1 name2
(the variants travel to the GPU box with the snapshot; bench.py picks them up through PCCM_LIB)."""
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "open_pcc_metric_b200", "csrc", "pccm_api.cu")


def so(name):
    return os.path.join(ROOT, "build", f"libpccm_{name}.so")


def build(specs):
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    procs = []
    for spec in specs:
        name, _, flags = spec.partition(":")
        cmd = [os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo",
               "--fmad=false", "-std=c++17", "-Xcompiler", "-fPIC,-O2", "-shared", "-o", so(name), SRC] + [f for f in flags.split(",") if f]
        procs.append((name, subprocess.Popen(cmd)))
    for name, p in procs:
        assert p.wait() == 0, name


def variant_env(name):
    """name[@VAR=VAL[@VAR=VAL...]]: library variant + environment switches"""
    lib, *sets = name.split("@")
    env = dict(os.environ, PCCM_LIB=so(lib) if lib != "base" else os.path.join(ROOT, "open_pcc_metric_b200", "libpccm.so"))
    for kv in sets:
        k, _, v = kv.partition("=")
        env[k] = v
    return env


def launch_list(name):
    env = variant_env(name)
    out = os.path.join(ROOT, "gpurun_out", f"launches_{name.replace('@', '_').replace('=', '')}.csv")
    subprocess.run(["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "-c", "400", "--csv", "--log-file", out,
                    sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "2", "--no-cpu-baseline", "--no-e2e"],
                   env=env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=False)
    rows = [r for r in csv.reader(open(out)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    names = [(r[ki].split("(")[0].replace("pccm::", ""), float(r[vi].replace(",", "")) / 1000) for r in rows[1:]]
    idx = [i for i, (n, _) in enumerate(names) if n.startswith("stats_kernel")]
    step = names[idx[-2]:]
    step = [x for x in step if "Fill" not in x[0]]
    return step


def run(names):
    res = {n: launch_list(n) for n in names}
    keys = [k for k, _ in res[names[0]]]
    print(f"{'kernel':28s}" + "".join(f"{n[-12:]:>13s}" for n in names))
    for n in names[1:]:
        if [k for k, _ in res[n]] != keys:      # different launch sequence: print it apart
            print(n, [(k, round(v, 2)) for k, v in res[n]])
    for i, k in enumerate(keys):
        print(f"{k:28s}" + "".join(f"{(res[n][i][1] if i < len(res[n]) and res[n][i][0] == k else float('nan')):13.2f}" for n in names))
    print(f"{'sum':28s}" + "".join(f"{sum(v for _, v in res[n]):13.2f}" for n in names))
    # un-profiled step time of each variant
    for n in names:
        env = variant_env(n)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "100", "--warmup", "5", "--no-cpu-baseline", "--no-e2e"],
                             env=env, capture_output=True, text=True).stdout.strip().splitlines()
        import json
        d = json.loads(out[-1])
        print(n, "ms_per_step", round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["stage_ms_per_step"].items() if v})


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        run(sys.argv[2:])
