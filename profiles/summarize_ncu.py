"""Summarise .ncu-rep captures (read on the CPU box: `ncu -i rep --page raw --csv`) into a small
markdown table of the metrics the roofline discussion needs.

    python profiles/summarize_ncu.py gpurun_out/r1_prof_gemm.ncu-rep [more.ncu-rep ...] > profiles/r1_ncu_summary.md
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor inst"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    header, units, data = rd[0], rd[1], rd[2:]
    return header, units, data


def main():
    for rep in sys.argv[1:]:
        header, units, data = rows_of(rep)
        idx = {h: i for i, h in enumerate(header)}
        print(f"### {rep}\n")
        cols = [k for k, _ in KEYS if k in idx]
        print("| kernel | " + " | ".join(dict(KEYS)[k] for k in cols) + " |")
        print("|---|" + "---|" * len(cols))
        for d in data:
            name = d[idx["Kernel Name"]]
            name = name.replace("fsmg::", "").replace("tc::", "")[:60]
            cells = []
            for k in cols:
                v, u = d[idx[k]], units[idx[k]]
                cells.append(f"{v} {u}".strip())
            print(f"| `{name}` | " + " | ".join(cells) + " |")
        print()


if __name__ == "__main__":
    main()
