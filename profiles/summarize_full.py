"""Key metrics per profiled launch from an `ncu --set full` report: either the .ncu-rep (read here with `ncu -i rep --page raw
--csv`) or that raw CSV already exported on the GPU box (reports with many launches exceed the 64 MiB that travel back)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct"]


def main(rep):
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    print(f"# {rep}")
    for r in rows[2:]:
        print("\n## " + r[kn][:100])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:68s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    main(sys.argv[1])
