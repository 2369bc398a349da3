#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/<tag>/prof_x.ncu-rep > profiles/<round>_<kernel>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_global_ld.sum",
    "smsp__sass_inst_executed_op_global_st.sum", "smsp__inst_executed_op_tma_ld.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    # --print-units base: every value in its base unit (bytes, ns, ...): columns of one report otherwise come out
    # scaled independently (dram read in Mbyte next to dram write in Gbyte)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: {len(data)} captured launch(es); ncu --set full --clock-control none (numbers under a profiler: evidence, not bench values)")
    for r in data:
        print(f"\n## {r[col['Kernel Name']][:110]}  grid={r[col['Grid Size']]} block={r[col['Block Size']]}")
        for k in KEYS:
            if k in col:
                print(f"{k:75s} {r[col[k]]:>16s} {units[col[k]]}")
        st = sorted(((float(r[i].replace(',', '') or 0), h[len(STALL):].replace('_per_issue_active.ratio', ''))
                     for h, i in col.items() if h.startswith(STALL) and h.endswith("_per_issue_active.ratio")), reverse=True)
        print("warp-stall reasons per issue-active cycle: " + ", ".join(f"{n}={v:.2f}" for v, n in st[:8]))
        try:
            def in_bytes(key):
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[units[col[key]]]
                return float(r[col[key]].replace(',', '')) * scale
            rd, wr = in_bytes("dram__bytes_read.sum"), in_bytes("dram__bytes_write.sum")
            print(f"traffic (dram read+write) = {(rd + wr) / 1e6:.3f} MB ({rd / 1e6:.3f} MB read + {wr / 1e6:.3f} MB written)")
        except Exception as e:
            print(f"traffic: not available ({e})")


if __name__ == "__main__":
    main()
