"""Condense an `ncu --page raw --csv` dump into the handful of rows the design notes cite.

    ncu -i X.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv "note" > profiles/X.summary.csv
"""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    print("# %s" % (sys.argv[2] if len(sys.argv) > 2 else ""))
    print("# kernel: %s" % d.get("Kernel Name", ("", "?"))[1])
    print("metric,unit,value")
    for k in KEYS:
        if k in d:
            print('%s,%s,%s' % (k, d[k][0], d[k][1].replace(",", "")))
    st = [(float(v.replace(",", "")), h) for h, (u, v) in d.items()
          if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    for v, h in sorted(st, reverse=True)[:8]:
        print("stall_%s,cycles per issued instruction,%.3f"
              % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))


main()
