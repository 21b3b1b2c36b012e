"""Print the handful of ncu metrics the roofline discussion needs from a .ncu-rep (read on the CPU box)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        print("---", row[hdr.index("Kernel Name")][:90], "grid", row[hdr.index("launch__grid_size")] if "launch__grid_size" in hdr else "")
        for w in WANT:
            if w in hdr:
                print(f"   {w:75s} {row[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
        for t in hdr:
            if t in ("sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
                     "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"):
                print(f"   {t:75s} {row[hdr.index(t)]:>16s} {units[hdr.index(t)]}")


if __name__ == "__main__":
    main(sys.argv[1])
