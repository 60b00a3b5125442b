"""Turns the scratch artefacts of a GPU run (gpurun_out/) into the committed summaries under profiles/.

  python scripts/make_profiles.py <tag>      e.g. r1_final

Reads (when present): gpurun_out/bench.log, ops.json, launches.csv, prof_igemm.ncu-rep, prof_fattn.ncu-rep,
matrix.jsonl, scale.jsonl, bench_convs.json, clocks.
"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.environ.get("GP_PROFILES_DIR") or os.path.join(ROOT, "profiles")   # on the GPU box: a directory under gpurun_out/


def cat(n):
    if "fattn" in n: return "fused attention (d=64)"
    if n.endswith("conv_in") and "encoder" in n: return "convolutions (igemm)"
    if ".softmax" in n: return "softmax rows (VAE d=512 attention)"
    if n.endswith(".qk") or n.endswith(".pv") or "to_vT" in n or "to_qk" in n: return "attention GEMMs (igemm)"
    if (".norm1" in n and "transformer" in n) or ".norm3" in n: return "LayerNorm"
    if "norm" in n: return "GroupNorm(+SiLU)"
    if "geglu" in n: return "FF projection + GEGLU (igemm)"
    if "attn2" in n: return "cross-attention closed form"
    if "ff." in n or "proj_" in n or "to_out" in n: return "linear layers (igemm)"
    return "convolutions (igemm)"


def ops_table(tag):
    f = os.path.join(G, "ops.json")
    if not os.path.exists(f):
        return ""
    ops = json.load(open(f))
    shutil.copy(f, os.path.join(P, f"{tag}_ops_per_op_events.json"))
    tot = sum(o["usec"] for o in ops)
    d = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
    for o in ops:
        c = cat(o["name"])
        d[c][0] += o["usec"]; d[c][1] += o["flops"]; d[c][2] += o["bytes"]; d[c][3] += 1
    out = [f"Per-op CUDA-event timing of one warm step (sum {tot / 1000:.1f} ms, {len(ops)} ops):", "",
           "| category | ms | share | ops | TFLOP/s (algorithmic) | GB/s (nominal bytes) |", "|---|---|---|---|---|---|"]
    for k, v in sorted(d.items(), key=lambda kv: -kv[1][0]):
        out.append(f"| {k} | {v[0] / 1000:.2f} | {100 * v[0] / tot:.1f}% | {v[3]} | {v[1] / max(v[0], 1e-9) / 1e6:.0f} | "
                   f"{v[2] / max(v[0], 1e-9) / 1e3:.0f} |")
    top = sorted(ops, key=lambda o: -o["usec"])[:12]
    out += ["", "Slowest ops:", "", "| us | TFLOP/s | op |", "|---|---|---|"]
    for o in top:
        out.append(f"| {o['usec']:.0f} | {o['flops'] / max(o['usec'], 1e-9) / 1e6:.0f} | `{o['name']}` |")
    return "\n".join(out)


def launches_table(tag):
    f = os.path.join(G, "launches.csv")
    if not os.path.exists(f):
        return ""
    lines = [l for l in open(f) if not l.startswith("==")]
    r = list(csv.reader(lines))
    idx = {h: i for i, h in enumerate(r[0])}
    tot = collections.defaultdict(lambda: [0, 0.0])
    for d in r[1:]:
        try:
            v = float(d[idx["Metric Value"]].replace(",", ""))
        except Exception:
            continue
        u = d[idx["Metric Unit"]]
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(u, 1)
        n = re.sub(r"\(.*", "", d[idx["Kernel Name"]]).replace("void gp::<unnamed>::", "").replace("gp::<unnamed>::", "")
        n = n.replace("void unnamed>::", "").replace("unnamed>::", "")
        tot[n][0] += 1; tot[n][1] += ns
    T = sum(v[1] for v in tot.values())
    out = ["ncu launch list of the bench command (`--metrics gpu__time_duration.sum --clock-control none`; cold-cache, "
           "serialised — compare SHARES):", "", "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {v[0]} | {v[1] / 1e6:.2f} | {100 * v[1] / T:.1f}% |")
    return "\n".join(out)


def step_dram(tag):
    """Per-launch DRAM traffic of ONE warm step (scripts/prof_step.py under ncu): totals by kernel, and the
    igemm total next to the algorithmic bytes of the same launches (ops.json)."""
    f = os.path.join(G, "step_dram.csv")
    if not os.path.exists(f):
        return ""
    lines = [l for l in open(f) if l.startswith('"')]
    r = list(csv.reader(lines))
    if len(r) < 10:
        return ""
    idx = {h: i for i, h in enumerate(r[0])}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3,
             "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
    tot = collections.defaultdict(lambda: [set(), 0.0, 0.0, 0.0])
    for d in r[1:]:
        try:
            v = float(d[idx["Metric Value"]].replace(",", "")) * scale.get(d[idx["Metric Unit"]], 1)
        except Exception:
            continue
        n = re.sub(r"\(.*", "", d[idx["Kernel Name"]]).replace("void unnamed>::", "").replace("unnamed>::", "")
        t = tot[n]
        t[0].add(d[idx["ID"]])
        m = d[idx["Metric Name"]]
        if m == "dram__bytes_read.sum": t[1] += v
        elif m == "dram__bytes_write.sum": t[2] += v
        elif m == "gpu__time_duration.sum": t[3] += v
    out = ["DRAM traffic of one warm step, every launch (`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`, "
           "`scripts/prof_step.py`):", "", "| kernel | launches | DRAM read GB | DRAM write GB | time ms (under ncu) |",
           "|---|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda kv: -(kv[1][1] + kv[1][2])):
        out.append(f"| `{k}` | {len(v[0])} | {v[1] / 1e9:.2f} | {v[2] / 1e9:.2f} | {v[3] * 1e3:.2f} |")
    ig = [v for k, v in tot.items() if "igemm" in k]
    res = {"igemm_launches": sum(len(v[0]) for v in ig), "igemm_dram_bytes": sum(v[1] + v[2] for v in ig),
           "all_dram_bytes": sum(v[1] + v[2] for v in tot.values())}
    of = os.path.join(G, "ops.json")
    if os.path.exists(of):
        ops = json.load(open(of))
        res["igemm_algorithmic_bytes"] = sum(o["bytes"] for o in ops if o.get("kind") == 1)
        res["all_nominal_bytes"] = sum(o["bytes"] for o in ops)
        out += ["", f"igemm launches: measured DRAM {res['igemm_dram_bytes'] / 1e9:.2f} GB per step vs "
                f"{res['igemm_algorithmic_bytes'] / 1e9:.2f} GB algorithmic (inputs + outputs + weights once per launch); "
                f"whole step {res['all_dram_bytes'] / 1e9:.2f} GB measured vs {res['all_nominal_bytes'] / 1e9:.2f} GB nominal."]
    json.dump(res, open(os.path.join(P, f"{tag}_step_dram.json"), "w"), indent=1)
    return "\n".join(out)


KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic"]


def ncu_extract(rep, tag, name, launches):
    f = os.path.join(G, rep)
    if not os.path.exists(f):
        return ""
    raw = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return ""
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = [f"`ncu --set full --clock-control none` extract ({rep}):", "", "```"]
    for rid, label in launches:
        if 2 + rid >= len(rows):
            continue
        r = rows[2 + rid]
        out.append(f"== launch {rid}: {label}")
        for k in KEYS:
            if k in idx:
                out.append(f"  {k:78s} {r[idx[k]][:48]:>48s} {units[idx[k]]}")
        out.append("")
    out.append("```")
    open(os.path.join(P, f"{tag}_{name}_ncu_set_full.txt"), "w").write("\n".join(out))
    return "\n".join(out)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(P, exist_ok=True)
    parts = []
    b = os.path.join(G, "bench.log")
    if os.path.exists(b):
        line = open(b).read().strip().splitlines()[-1]
        open(os.path.join(P, f"{tag}_bench_line.json"), "w").write(line + "\n")
        j = json.loads(line)
        parts.append(f"Bench line (`python bench.py`): **{j['value']:.1f} {j['unit']}** device-resident, "
                     f"**{j['e2e']['value']:.1f}** end to end (host buffers), {j['ms_per_step']:.1f} ms / step of "
                     f"{j['config']['global_batch']} images; igemm {j['roofline']['achieved']:.0f} TFLOP/s = "
                     f"{100 * j['roofline']['frac']:.0f}% of the measured sustained bf16 peak; clocks {j['clocks']}.")
    for name in ("matrix.jsonl", "scale.jsonl", "bench_convs.json"):
        if os.path.exists(os.path.join(G, name)):
            shutil.copy(os.path.join(G, name), os.path.join(P, f"{tag}_{name}"))
    parts.append(ops_table(tag))
    parts.append(launches_table(tag))
    parts.append(step_dram(tag))
    parts.append(ncu_extract("prof_igemm.ncu-rep", tag, "igemm",
                             [(3, "conv3x3 128->128 @768x768 B=8 (patch-resident main loop)"),
                              (7, "conv3x3 256->256 @384x384 B=8 (tap-streaming main loop, BN=256)"),
                              (11, "linear 320->2560 on 73728 tokens"),
                              (15, "K-packed stem: 1x1 GEMM 32(27)->128 on 8 x 768 x 768 pixels"),
                              (16, "conv3x3 128->128 @768x768 B=8 with GroupNorm+SiLU in the operand path (GP_GN_FUSE=1)")]))
    parts.append(ncu_extract("prof_fattn.ncu-rep", tag, "fattn", [(1, "fused attention T=9216, 5 heads, d=64, B=8")]))
    open(os.path.join(P, f"{tag}_summary.md"), "w").write("\n\n".join(p for p in parts if p) + "\n")
    print("\n\n".join(p for p in parts if p)[:3000])


if __name__ == "__main__":
    main()
