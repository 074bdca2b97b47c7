"""Summarise an .ncu-rep (ncu --set full) into the handful of lines DESIGN.md and
bench.py cite.  Usage: python profiles/summarize.py gpurun_out/prof_k1.ncu-rep > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
        "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
        "smsp__inst_executed_op_tma_ld.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


def main(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
        print("kernel:", d.get("Kernel Name", ("?",))[0])
        for k in KEYS:
            if k in d:
                print("  %-72s %s %s" % (k, d[k][0], d[k][1]))
        print("  -- warp stall reasons (per issue-active cycle)")
        st = [(float(v[0]), h) for h, v in d.items() if "issue_stalled" in h and h.endswith(".ratio")]
        for v, h in sorted(st, reverse=True)[:8]:
            print("  %-72s %.3f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace(
                "_per_issue_active.ratio", ""), v))
        rd = float(d["dram__bytes_read.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_read.sum"][1]]
        wr = float(d["dram__bytes_write.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[d["dram__bytes_write.sum"][1]]
        print("  dram_bytes_per_launch %d" % (rd + wr))


if __name__ == "__main__":
    main(sys.argv[1])
