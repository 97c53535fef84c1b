"""Turns ncu outputs brought back in gpurun_out/ into the small CSVs committed under profiles/.

  python scripts/ncu_summarize.py launches gpurun_out/launches_r01c.csv profiles/r01_launches_summary.csv
  python scripts/ncu_summarize.py full gpurun_out/prof_small_r01.ncu-rep profiles/r01_small_kernels_ncu_full_summary.csv
"""
import csv, io, re, subprocess, sys
from collections import OrderedDict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("<unnamed>::", "")
    return name.strip()


def launches(src, dst):
    rows = [l for l in open(src) if not l.startswith("==")]
    agg = OrderedDict()
    for r in csv.DictReader(io.StringIO("".join(rows))):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1e-3)
        a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("kernel,launches,total_us,share,avg_us\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{n},{t:.1f},{t / tot:.4f},{t / n:.2f}\n")
        f.write(f"TOTAL,{sum(a[0] for a in agg.values())},{tot:.1f},1.0,\n")


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    head, units, body = rd[0], rd[1], rd[2:]
    col = {h: i for i, h in enumerate(head)}
    keep = [m for m in METRICS if m in col]
    with open(dst, "w") as f:
        f.write("launch,kernel,grid,block," + ",".join(f"{m} [{units[col[m]]}]" for m in keep) + "\n")
        for i, r in enumerate(body):
            f.write(f"{i},\"{short(r[col['Kernel Name']])}\",\"{r[col['Grid Size']]}\",\"{r[col['Block Size']]}\"," +
                    ",".join(r[col[m]].replace(",", "") for m in keep) + "\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
