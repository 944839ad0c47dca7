"""Summarise .ncu-rep captures (gpurun_out/*.ncu-rep) into profiles/<round>_ncu_summary.{csv,md}.
usage: python tools/ncu_summary.py r01 gpurun_out/prof_a.ncu-rep [...]"""
import csv, io, os, subprocess, sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max.per_second", "smsp__inst_executed.sum",
]

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return [{h: (v, u) for h, u, v in zip(hdr, units, vals)} for vals in rows[2:]]

def main():
    tag, reps = sys.argv[1], sys.argv[2:]
    os.makedirs("profiles", exist_ok=True)
    lines = ["| capture | kernel | " + " | ".join(METRICS) + " |", "|" + "---|" * (len(METRICS) + 2)]
    table = []
    for rep in reps:
        for d in raw(rep):
            name = d.get("Kernel Name", ("?", ""))[0]
            row = {"capture": os.path.basename(rep), "kernel": name}
            for m in METRICS:
                v, u = d.get(m, ("", ""))
                row[m] = f"{v} {u}".strip()
            table.append(row)
            lines.append("| " + row["capture"] + " | " + name[:70] + " | " + " | ".join(row[m] for m in METRICS) + " |")
    with open(f"profiles/{tag}_ncu_summary.csv", "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=["capture", "kernel"] + METRICS)
        w.writeheader(); w.writerows(table)
    with open(f"profiles/{tag}_ncu_summary.md", "w") as f:
        f.write("\n".join(lines) + "\n")
    print(f"wrote profiles/{tag}_ncu_summary.csv/.md with {len(table)} kernels")

if __name__ == "__main__":
    main()
