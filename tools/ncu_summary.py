"""Summarise an .ncu-rep (read here, no GPU): python tools/ncu_summary.py gpurun_out/x.ncu-rep [pattern ...]"""
import csv, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
        "gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct"]


def main():
    rep = sys.argv[1]
    pats = sys.argv[2:]
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("==", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
        for i, h in enumerate(hdr):
            short = h.split(".", 2)[-1] if h.startswith(("SM_A.", "TPC.")) else h
            if short in KEYS or any(p in h for p in pats):
                print(f"  {h} [{units[i]}] = {vals[i]}")


if __name__ == "__main__":
    main()
