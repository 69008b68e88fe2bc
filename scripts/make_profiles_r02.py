"""Turns the grouped ncu captures of scripts/gpu_r2_ncu.sh (gpurun_out/<tag>_<group>.csv) into the committed summaries
profiles/r02_<tag>_ncu_metrics.txt and profiles/ncu_traffic.json (read by bench.py for roofline.traffic), and copies the
launch lists of the default bench commands."""
import csv
import glob
import json
import os
import re
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
KERNELS = {"wn6": ("wavenet6_kernel", "fp32 layer-pipelined WaveNet (csrc/wavenet6.cu)"),
           "wn7": ("wavenet7_kernel", "bf16 tcgen05 layer-pipelined WaveNet (csrc/wavenet7.cu)"),
           "sr": ("samplernn_cluster_kernel", "SampleRNN cluster kernel, lane-major fp32 frame-tier engine (csrc/samplernn2.cu, ENGINE 1)"),
           "srtc": ("samplernn_cluster_kernel", "SampleRNN cluster kernel, tcgen05 bf16 frame-tier engine (csrc/samplernn2.cu, ENGINE 2)")}


def rows(path, kname):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    out = []
    for r in csv.DictReader(lines):
        if kname.split("_kernel")[0] in r.get("Kernel Name", ""):
            out.append((r["Metric Name"], r["Metric Unit"], r["Metric Value"], r.get("Kernel Name", ""), r.get("Grid Size", ""), r.get("Block Size", "")))
    return out


traffic = {}
for tag, (kname, what) in KERNELS.items():
    files = sorted(glob.glob(os.path.join(OUT, f"{tag}_*.csv")))
    if not files:
        continue
    cfg = None
    log = os.path.join(OUT, f"ncu_dram_{tag}.log")
    if os.path.exists(log):
        m = re.search(r'"config": (\{[^}]*\})', open(log).read())
        cfg = json.loads(m.group(1)) if m else None
    vals = {}
    with open(os.path.join(PROF, f"r02_{tag}_ncu_metrics.txt"), "w") as f:
        f.write(f"# {what}\n# ncu --clock-control none --metrics <group> -k regex:{kname.split('_kernel')[0]} -s 1 -c 1 python bench.py ... "
                f"(scripts/gpu_r2_ncu.sh; one run per metric group: --set full dies with LaunchFailed on the layer-pipelined kernels)\n")
        if cfg:
            f.write(f"# captured launch: {cfg['workload']}\n")
        for path in files:
            f.write(f"\n[{os.path.basename(path)[len(tag) + 1:-4]}]\n")
            for name, unit, val, kn, grid, block in rows(path, kname):
                f.write(f"{name:95s} {unit:8s} {val}\n")
                vals[name] = float(val.replace(",", ""))
                kinfo = (kn, grid, block)
        f.write(f"\n# kernel {kinfo[0][:80]} grid {kinfo[1]} block {kinfo[2]}\n")
    if "dram__bytes_read.sum" in vals and cfg:
        traffic[kname + ("<2>" if tag == "srtc" else "")] = {"dram_bytes": vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"],
                          "l2_bytes": vals.get("lts__t_bytes.sum"), "duration_ns": vals.get("gpu__time_duration.sum"),
                          "prompts": cfg["batch_per_gpu"], "prompt_len": cfg["prompt_len"], "n_steps": cfg["n_steps"],
                          "source": f"profiles/r02_{tag}_ncu_metrics.txt (dram__bytes_read.sum + dram__bytes_write.sum of one launch: "
                                    f"{cfg['batch_per_gpu']} prompts x ({cfg['prompt_len']} prompt + {cfg['n_steps']} generated samples); "
                                    "bench.py scales it by prompts x (prompt + generated) samples"}
with open(os.path.join(PROF, "ncu_traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1)
for name in ("wavenet", "samplernn", "features"):
    src = os.path.join(OUT, f"r02_{name}_launches.csv")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(PROF, f"r02_{name}_launches.csv"))
print(json.dumps(traffic, indent=1))
