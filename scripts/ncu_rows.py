"""Key rows of an `ncu --set full` report: python scripts/ncu_rows.py REPORT.ncu-rep  (reads `ncu -i ... --page raw --csv`)."""
import csv
import io
import subprocess
import sys

WANT = ['Kernel Name', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active']


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    rows = [r for r in rows if len(r) > 10]
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(w) for w in WANT if w in hdr]
    print("| " + " | ".join(f"{hdr[i]} [{units[i]}]" for i in idx) + " |")
    print("|" + "---|" * len(idx))
    for r in rows[2:]:
        print("| " + " | ".join((r[i].replace("hn::", "")[:70] if hdr[i] == 'Kernel Name' else r[i]) for i in idx) + " |")


if __name__ == "__main__":
    main()
