"""Summaries of the ncu artefacts gpurun brought back (gpurun_out/) -> profiles/ (tracked).
usage: python scripts/summarize_ncu.py TAG [--launches FILE.csv] [REP ...]   (TAG like r02; REP = name of gpurun_out/REP.ncu-rep)"""
import collections
import csv
import subprocess
import sys

tag = sys.argv[1]
args = sys.argv[2:]
launches = "gpurun_out/launches.csv"
if args and args[0] == "--launches":
    launches = args[1]; args = args[2:]
rows = list(csv.reader(open(launches)))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            v = float(d["Metric Value"].replace(",", ""))
            ns = v * {"ns": 1, "us": 1e3, "ms": 1e6}.get(d["Metric Unit"], 1)
            k = d["Kernel Name"].split("(")[0].replace("void ", "")
            agg[k][0] += 1
            agg[k][1] += ns
tot = sum(v[1] for v in agg.values())
with open(f"profiles/{tag}_launches_summary.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 1 --no-cpu-baseline\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
    f.write(f"{'kernel':42s} {'launches':>8s} {'total_ms':>10s} {'share_%':>8s} {'avg_us':>10s}\n")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{k[:42]:42s} {v[0]:8d} {v[1] / 1e6:10.3f} {v[1] / tot * 100:8.1f} {v[1] / v[0] / 1e3:10.1f}\n")
open(f"profiles/{tag}_launches.csv", "w").write(open(launches).read())

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for rep in args:
    out = subprocess.run(["ncu", "-i", f"gpurun_out/{rep}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(out.splitlines()))
    h = rr[0]
    idx = [h.index(w) for w in WANT if w in h]
    name = rep[len(tag) + 1:] if rep.startswith(tag + "_") else rep
    with open(f"profiles/{tag}_{name}_full.csv", "w") as f:
        w = csv.writer(f)
        w.writerow([h[i] for i in idx]); w.writerow([rr[1][i] for i in idx])
        for r in rr[2:]:
            w.writerow([r[i] for i in idx])
    print(open(f"profiles/{tag}_{name}_full.csv").read())
print(open(f"profiles/{tag}_launches_summary.txt").read())
