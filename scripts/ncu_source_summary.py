"""Source-page summary of one launch of an .ncu-rep (captured with --set full --import-source on): headline metrics,
warp instructions per execution-count class (= per tile / per piece / per K block ...), opcode histogram of the hottest
class, stall-reason totals and the most-sampled instructions.

  python scripts/ncu_source_summary.py gpurun_out/prof_linear.ncu-rep [launch_index] [tiles] > profiles/<tag>_..._ncu_source.txt
"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tiles = int(sys.argv[3]) if len(sys.argv) > 3 else 0


def ncu(*a):
    return subprocess.run(["ncu", "-i", rep, *a], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units, row = raw[0], raw[1], raw[2 + launch]
idx = {h: i for i, h in enumerate(hdr)}
print(f"{rep}, launch {launch}: {row[idx['Kernel Name']]}")
for m in ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
          "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread"]:
    if m in idx:
        print(f"  {m:75s} {row[idx[m]]} {units[idx[m]]}")

src = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--launch-skip", str(launch), "--launch-count", "1"))))
h = src[1]
ix = {n: i for i, n in enumerate(h)}
data = src[2:]
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_samp = sum(int(r[ix["# Samples"]]) for r in data)
print(f"\nwarp instructions executed: {tot_inst}" + (f" = {tot_inst / tiles:.0f} per tile ({tiles} tiles)" if tiles else ""))
cls = collections.Counter()
smp = collections.Counter()
for r in data:
    e = int(r[ix["Instructions Executed"]])
    cls[e] += 1
    smp[e] += int(r[ix["# Samples"]])
print("\nexecution-count classes (executions of each instruction, number of such SASS instructions, warp instructions, samples):")
for e, n in sorted(cls.items(), key=lambda kv: -kv[0] * kv[1])[:10]:
    per = f"  ({e / tiles:.1f} per tile)" if tiles else ""
    print(f"  {e:9d} x {n:4d} = {e * n:10d}   samples {smp[e]:6d}{per}")
hot = max(cls.items(), key=lambda kv: kv[0] * kv[1])[0]
ops = collections.Counter()
for r in data:
    if int(r[ix["Instructions Executed"]]) == hot:
        t = r[ix["Source"]].split()
        ops[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += 1
print(f"\nopcodes of the hottest class ({hot} executions each): " + ", ".join(f"{k} {v}" for k, v in ops.most_common(16)))
agg = {n: sum(int(r[ix[n]]) for r in data) for n in stalls}
print(f"\nstall samples ({tot_samp} total): " + ", ".join(f"{k[6:]} {100 * v / max(tot_samp, 1):.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
print("\nmost-sampled instructions:")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:14]:
    s = int(r[ix["# Samples"]])
    top = sorted(((int(r[ix[n]]), n[6:]) for n in stalls), reverse=True)[0]
    print(f"  {100 * s / max(tot_samp, 1):4.1f}%  {r[ix['Source']].strip()[:60]:60s} x{r[ix['Instructions Executed']]:>8s}  {top[1]}")
