"""Per-launch summary of an .ncu-rep (read on the CPU box): duration, DRAM bytes and throughput,
achieved occupancy, tensor-pipe share.  Usage: python tools/ncu_summary.py file.ncu-rep [name-regex]"""
import csv
import io
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(v.replace(",", "")) * scale


def to_us(v, unit):
    scale = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)
    return float(v.replace(",", "")) * scale


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    print(f"{'kernel':58s} {'us':>9s} {'rd MB':>9s} {'wr MB':>9s} {'GB/s':>8s} {'dram%':>6s} {'sm%':>6s} {'issue%':>6s} {'occ%':>6s} {'regs':>4s} {'Minst':>7s}")
    for r in data:
        name = r[col["Kernel Name"]]
        if pat and not pat.search(name):
            continue
        g = lambda m: (r[col[m]], units[col[m]]) if m in col else ("0", "")
        us = to_us(*g("gpu__time_duration.sum"))
        rd, wr = to_bytes(*g("dram__bytes_read.sum")), to_bytes(*g("dram__bytes_write.sum"))
        f = lambda m: float(g(m)[0].replace(",", "") or 0) if m in col else float("nan")
        short = re.sub(r"\(.*", "", name).replace("<unnamed>::", "")
        print(f"{short[:58]:58s} {us:9.1f} {rd / 1e6:9.2f} {wr / 1e6:9.2f} {(rd + wr) / us / 1e3:8.0f} "
              f"{f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} {f('sm__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
              f"{f('smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} "
              f"{f('sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} {int(f('launch__registers_per_thread')):4d} "
              f"{f('smsp__inst_executed.sum') / 1e6:7.1f}")


if __name__ == "__main__":
    main()
