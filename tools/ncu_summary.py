"""Summarise an .ncu-rep (read with `ncu -i`) into a small text file for profiles/."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_imma", "sm__inst_executed_pipe_tensor_subpipe_hmma", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__pcsamp_warps_issue_stalled", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed_pipe_fma", "smsp__inst_executed_pipe_alu", "sm__inst_executed_pipe_fmaheavy",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "local_load", "local_store", "lts__t_bytes.sum"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H = rows[0]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of {rep}\n")
        for i, h in enumerate(H):
            if h == "Kernel Name" or any(k in h for k in KEYS):
                if "_not_issued" in h:
                    continue
                f.write(h + " | " + " | ".join(r[i][:48] for r in rows[1:]) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
