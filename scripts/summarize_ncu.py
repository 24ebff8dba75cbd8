#!/usr/bin/env python3
"""Turns an ncu report (.ncu-rep, brought back in gpurun_out/) into the tracked summaries under
profiles/: one markdown table per capture with duration, DRAM traffic, pipe utilisation, occupancy
and the PC-sampling stall mix of every kernel of this repo.

  python scripts/summarize_ncu.py gpurun_out/prof_r01.ncu-rep profiles/ncu_r01_full.md [songs [traffic.json]]

With `traffic.json` it also writes the per-kernel DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per
launch and per song that bench.py reports as `roofline.traffic`, keyed by the kernel names of the bench JSON.
"""
import csv
import io
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    songs = int(sys.argv[3]) if len(sys.argv) > 3 else None
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, data = rows[0], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def g(r, name, default=""):
        return r[col[name]] if name in col else default

    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    lines = ["# ncu --set full summary of `%s`" % rep, ""]
    if songs:
        lines.append("Workload: bench.py step over %d synthetic 3-min tracks (one launch of each kernel)." % songs)
        lines.append("")
    lines.append("| kernel | grid x block | regs | time ms | DRAM rd GB | DRAM wr GB | DRAM % peak | issue active % | "
                 "warps active % | fma pipe % | fp64 pipe % | smem wavefronts M | bank conflicts M | top stalls |")
    lines.append("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for r in data:
        name = g(r, "Kernel Name").split("(")[0].replace("void ", "")
        tot = sum(float(r[i] or 0) for i in stall_cols) or 1.0
        st = sorted(((float(r[i] or 0) / tot * 100, hdr[i].replace("smsp__pcsamp_warps_issue_stalled_", "")) for i in stall_cols), reverse=True)[:4]
        f = lambda k, d=2: ("%%.%df" % d) % float(g(r, k, "0") or 0)
        lines.append("| %s | %s x %s | %s | %s | %s | %s | %s | %s | %s | %s | %s | %.1f | %.1f | %s |" % (
            name, g(r, "launch__grid_size"), g(r, "launch__block_size"), g(r, "launch__registers_per_thread"),
            f("gpu__time_duration.sum", 3), f("dram__bytes_read.sum", 3), f("dram__bytes_write.sum", 3),
            f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1), f("smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
            f("sm__warps_active.avg.pct_of_peak_sustained_active", 1), f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1),
            f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 1),
            float(g(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "0") or 0) / 1e6,
            float(g(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "0") or 0) / 1e6,
            ", ".join("%s %.0f%%" % (n, v) for v, n in st)))
    lines.append("")
    lines.append("Units as printed by `ncu --page raw --csv` (ms, GB = 1e9 bytes). Captured with "
                 "`--set full --clock-control none --import-source on`; durations under ncu are serialised and "
                 "cold-cache, compare shares, not absolutes (the timed numbers are in the bench JSON).")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)
    if len(sys.argv) > 4 and songs:
        import json
        import re
        # bench.py's kernel ids (api.cu kKernelNames) for the current kernels
        alias = {"chroma_pipe_kernel": "chroma_kernel", "tuning_select_kernel": "tuning_kernel",
                 "stft8192v2_kernel": "stft8192_kernel", "stft8192v3_kernel": "stft8192_kernel",
                 "pvoc512v2_kernel": "pvoc512_kernel", "distance_matrix_diag_kernel": "distance_kernels"}
        tr = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from ncu --set full (%s), one launch = %d "
                          "synthetic 3-min songs; bench.py scales bytes_per_song by the songs per launch" % (out, songs),
              "songs_per_launch": songs}
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        units = rows[1]
        for r in data:
            name = re.sub(r"<.*", "", g(r, "Kernel Name").split("(")[0].replace("void ", "").replace("bliss::", ""))
            name = alias.get(name, name)
            b = sum(float(g(r, k, "0") or 0) * unit.get(units[col[k]], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            pct = lambda k: round(float(g(r, k, "0") or 0), 1)
            tr[name] = {"bytes_per_launch": b, "bytes_per_song": b / songs,
                        # where the kernel actually sits (same capture): the unified L1 / shared-memory data pipe,
                        # issue slots, DRAM, L2
                        "l1tex_data_pipe_pct": pct("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                        "issue_active_pct": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                        "fma_pipe_pct": pct("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                        "dram_pct": pct("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                        "lts_pct": pct("lts__throughput.avg.pct_of_peak_sustained_elapsed")}
        json.dump(tr, open(sys.argv[4], "w"), indent=1)
        print("wrote", sys.argv[4])


if __name__ == "__main__":
    main()
