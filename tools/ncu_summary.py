"""Condense an `ncu --set full` report into the handful of numbers DESIGN.md / profiles/ quote.
    python tools/ncu_summary.py gpurun_out/conv_r1b.ncu-rep [--md profiles/x.md]
Runs `ncu -i <rep> --page raw --csv` (works without a GPU) and prints one row per captured launch.
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "us"),
    ("sm__cycles_elapsed.max", "cyc"),
    ("smsp__cycles_active.avg", "act cyc"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor %"),
    ("sm__inst_executed_pipe_uniform.sum", "uni inst"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem lsu %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM MB"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "L2->SM TB/s"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("dram__bytes_read.sum", "dram rd MB"),
    ("dram__bytes_write.sum", "dram wr MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
    ("launch__registers_per_thread", "regs"),
    ("lts__t_requests_srcunit_tex_op_red.sum", "L2 red req"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st long_sb"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st barrier"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "st membar"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st lg_thr"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st short_sb"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    names = ["kernel", "grid", "block"] + [f"{lab}" for _, lab in KEYS]
    out.append("| " + " | ".join(names) + " |")
    out.append("|" + "---|" * len(names))
    for d in data:
        cells = [d[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("s3d::", ""), d[idx["Grid Size"]], d[idx["Block Size"]]]
        for k, _ in KEYS:
            if k in idx:
                v = d[idx[k]]
                try:
                    f = float(v.replace(",", ""))
                    u = units[idx[k]]
                    if k.endswith("bytes.sum") or k.endswith("read.sum") or k.endswith("write.sum"):
                        f = f / 1e6 if u == "byte" else (f / 1e3 if u == "Kbyte" else (f * 1e3 if u == "Gbyte" else f))
                    v = f"{f:.3g}" if abs(f) < 1e6 else f"{f:.4g}"
                except ValueError:
                    pass
                cells.append(v)
            else:
                cells.append("-")
        out.append("| " + " | ".join(cells) + " |")
    text = "\n".join(out)
    if "--md" in sys.argv:
        open(sys.argv[sys.argv.index("--md") + 1], "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
