"""Compact summary of an `ncu --set full` report: the metrics the roofline discussion needs
(duration, DRAM bytes, tensor-pipe activity, L2 / SM throughput, registers, local-memory traffic).
Usage: python tools/ncu_summary.py report.ncu-rep [algorithmic_flop] [algorithmic_bytes] > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "gpc__cycles_elapsed.max.per_second", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
]
SCALE = {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("kernel:", d.get("Kernel Name"), " grid", d.get("Grid Size"), " block", d.get("Block Size"))
    vals = {}
    for k in KEYS:
        if k in d:
            u = units[hdr.index(k)]
            print(f"  {k:78s} {d[k]:>16s} {u}")
            vals[k] = (d[k], u)

    def num(k):
        v, u = vals[k]
        return float(v.replace(",", "")) * SCALE.get(u, 1.0)
    try:
        t = num("gpu__time_duration.sum")
        by = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        print(f"  -> DRAM traffic {by / 1e6:.1f} MB = {by / t / 1e9:.0f} GB/s")
        if len(sys.argv) > 2:
            print(f"  -> algorithmic {float(sys.argv[2]) / 1e9:.2f} GFLOP = {float(sys.argv[2]) / t / 1e12:.0f} TFLOP/s")
        if len(sys.argv) > 3:
            print(f"  -> algorithmic bytes {float(sys.argv[3]) / 1e6:.1f} MB (traffic / algorithmic = {by / float(sys.argv[3]):.2f})")
    except Exception as e:  # noqa: BLE001
        print("  (derived values unavailable:", e, ")")
