"""Key metrics per captured launch from an .ncu-rep (`ncu -i rep --page raw --csv`), as text for profiles/.

usage: python tools/ncu_rep_summary.py gpurun_out/prof_convfwd.ncu-rep > profiles/<name>_summary.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpc__cycles_elapsed.avg.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum", "smsp__cycles_active.avg"]


def main(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        print("----")
        for k in KEYS:
            if k in head:
                i = head.index(k)
                print(f"{k} [{units[i]}] = {r[i][:110]}")


if __name__ == "__main__":
    main(sys.argv[1])
