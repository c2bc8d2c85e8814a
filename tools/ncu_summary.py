"""Turn an `ncu --set full` report into the per-kernel table committed under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_all.ncu-rep profiles/r1_kernels_ncu.md [--title "..."]
    python tools/ncu_summary.py gpurun_out/prof_all_raw.csv profiles/r1_kernels_ncu.md      # exported raw page

Runs `ncu -i <rep> --page raw --csv` (works without a GPU) and keeps, per kernel launch: duration, the
integer / memory pipe utilisations, DRAM bytes and bandwidth against the measured HBM peak, occupancy,
registers, shared memory and the top warp-stall reason.
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU(POPC) %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("dram__bytes_read.sum", "DRAM rd"),
    ("dram__bytes_write.sum", "DRAM wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % (ncu)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("launch__shared_mem_per_block_static", "static smem"),
]

UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6,
              "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0, "nsecond": 1e-9}


def short_name(k: str) -> str:
    k = re.sub(r"\(.*$", "", k)
    k = k.replace("plm::", "").replace("void ", "").replace("(anonymous namespace)::", "")
    return k.strip()


def fnum(x: str) -> float:
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[sys.argv.index("--title") + 1] if "--title" in sys.argv else os.path.basename(rep)
    if rep.endswith(".csv"):  # the raw page already exported on the GPU box (the .ncu-rep itself can exceed gpurun's 64 MiB)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm_peak = float(peaks["hbm_gbs"])
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
                  and "selected" not in h]
    lines = [f"# {title}", "",
             "`ncu --set full --clock-control none --import-source on` over `python tools/profile_kernels.py` "
             "(the same command exited 0 without ncu first). Durations are ncu's serialised, cold-cache figures; "
             f"DRAM GB/s = (read + write bytes) / duration against the measured HBM copy peak {hbm_peak:.0f} GB/s "
             "(MEASURED_PEAKS.json).", "",
             "| # | kernel | grid x block | " + " | ".join(n for _, n in WANT) + " | DRAM GB/s | % HBM peak | top stall |",
             "|---|---|---|" + "---|" * (len(WANT) + 3)]
    for r in data:
        if len(r) < len(hdr):
            continue
        name = short_name(r[col["Kernel Name"]])
        if name.startswith("at::"):  # torch's own kernels (input generation), not part of the library
            continue
        cells = []
        t_s = rd = wr = float("nan")
        for m, _ in WANT:
            if m not in col:
                cells.append("-")
                continue
            v, u = r[col[m]], units[col[m]]
            x = fnum(v)
            if m == "gpu__time_duration.sum":
                t_s = x * UNIT_SCALE.get(u, 1.0)
                cells.append(f"{t_s * 1e6:.1f} us")
            elif m.startswith("dram__bytes"):
                b = x * UNIT_SCALE.get(u, 1.0)
                if "read" in m:
                    rd = b
                else:
                    wr = b
                cells.append(f"{b / 1e6:.3f} MB")
            elif "shared_mem" in m:
                cells.append(f"{x * UNIT_SCALE.get(u.split('/')[0], 1.0) / 1e3:.1f} KB")
            elif m == "launch__registers_per_thread":
                cells.append(f"{int(x)}")
            else:
                cells.append(f"{x:.1f}")
        gbs = (rd + wr) / t_s / 1e9 if t_s == t_s and t_s > 0 else float("nan")
        stalls = sorted(((fnum(r[col[h]]), h) for h in stall_cols), reverse=True)
        top = stalls[0][1].replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "") if stalls else "-"
        lines.append(f"| {r[col['ID']]} | `{name}` | {r[col['Grid Size']]} x {r[col['Block Size']]} | " + " | ".join(cells) +
                     f" | {gbs:.1f} | {100 * gbs / hbm_peak:.2f} | {top} ({stalls[0][0]:.2f}) |")
    lines.append("")
    open(out, "w").write("\n".join(lines))
    print(f"wrote {out}: {len(data)} launches")


if __name__ == "__main__":
    main()
