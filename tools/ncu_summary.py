"""Print the metrics we steer by from an .ncu-rep (run where ncu is installed; no GPU needed).

    python tools/ncu_summary.py gpurun_out/prof_k_first_order.ncu-rep
"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_bytes.sum", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("== %s  grid %s block %s" % (d.get("Kernel Name", "?")[:60], d.get("launch__grid_size"), d.get("launch__block_size")))
        for w in WANT:
            if w in d:
                print("  %-72s %s %s" % (w, d[w], units[hdr.index(w)]))
        stalls = sorted(((float(v.replace(",", "")), k) for k, v in d.items()
                         if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and v),
                        reverse=True)
        for v, k in stalls[:8]:
            print("  stall %-60s %.3f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
