"""Print selected metrics of an .ncu-rep (raw page) per kernel launch."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]


def main(path, extra=()):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    for r in rows[2:]:
        print("----", r[hdr.index("Kernel Name")][:60], r[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
        for w in list(WANT) + list(extra):
            if w in hdr:
                print(f"  {w:88s} {r[hdr.index(w)]}")
        st = sorted(((float(r[hdr.index(h)].replace(',', '') or 0), h) for h in stall), reverse=True)[:6]
        for v, h in st:
            print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:40s} {v:.2f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
