#!/usr/bin/env python3
"""Summarise an `ncu --set full` report into the few lines the roofline arithmetic needs.

usage: python tools/ncu_raw_summary.py gpurun_out/prof.ncu-rep [algorithmic_bytes] [flops] > profiles/rNN_ncu_<what>.txt

Reads `ncu -i <rep> --page raw --csv` (no GPU needed) and prints, per captured launch: duration, SM
clock, DRAM bytes read / written (-> `roofline.traffic`), L2->SM bytes, tensor-pipe activity, issue
activity, registers, grid — and, given the algorithmic bytes / FLOPs of the launch, the ratios.
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.avg", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__cluster_size", "smsp__inst_executed.sum",
]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9,
        "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def main():
    rep = sys.argv[1]
    algo_bytes = float(sys.argv[2]) if len(sys.argv) > 2 else None
    flops = float(sys.argv[3]) if len(sys.argv) > 3 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        unit = dict(zip(hdr, units))
        print(f"kernel: {rec.get('Kernel Name', '?')[:110]}")
        vals = {}
        for name in WANT:
            if name in rec and rec[name] != "":
                print(f"  {name:72s} {rec[name]:>16s} {unit.get(name, '')}")
                try:
                    vals[name] = float(rec[name].replace(",", "")) * UNIT.get(unit.get(name, ""), 1.0)
                except ValueError:
                    pass
        dur = vals.get("gpu__time_duration.sum")
        rd, wr = vals.get("dram__bytes_read.sum"), vals.get("dram__bytes_write.sum")
        if dur and rd is not None and wr is not None:
            print(f"  -> traffic (dram read + write) = {rd + wr:.0f} B, {(rd + wr) / dur / 1e9:.0f} GB/s under ncu")
            if algo_bytes:
                print(f"  -> traffic / algorithmic bytes = {(rd + wr) / algo_bytes:.4f}")
        xbar = vals.get("l1tex__m_xbar2l1tex_read_bytes.sum")
        if xbar and algo_bytes:
            print(f"  -> L2->SM bytes / algorithmic bytes = {xbar / algo_bytes:.3f}")
        if dur and flops:
            print(f"  -> {flops / dur / 1e12:.0f} TFLOP/s under ncu (a number under a profiler is not a bench value)")
        print()


if __name__ == "__main__":
    main()
