"""Summarise an ``.ncu-rep`` (``ncu --set full``) into one CSV row per kernel launch with the metrics the
profiling recipe names (duration, DRAM bytes / throughput, L2, SM, tensor pipe, registers, occupancy,
NVLink rx/tx when present).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [more.ncu-rep ...] > profiles/ncu/summary.csv
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "nvlrx__bytes.sum", "nvltx__bytes.sum", "pcie__read_bytes.sum", "pcie__write_bytes.sum",
]


def rows_of(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = csv.reader(io.StringIO(out))
    header = next(rd)
    next(rd, None)  # units row
    idx = {h: i for i, h in enumerate(header)}
    name_i = idx.get("Kernel Name")
    for r in rd:
        if not r or name_i is None or len(r) <= name_i:
            continue
        yield [path.split("/")[-1], r[name_i]] + [r[idx[m]] if m in idx and idx[m] < len(r) else "" for m in METRICS]


def main():
    w = csv.writer(sys.stdout)
    w.writerow(["report", "kernel"] + METRICS)
    for p in sys.argv[1:]:
        for row in rows_of(p):
            w.writerow(row)


if __name__ == "__main__":
    main()
