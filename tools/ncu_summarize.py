"""Turn a .ncu-rep (one kernel, `ncu --set full --import-source on`) into the two summaries kept under profiles/:
  <out>_raw.json            selected raw metrics of the launch (time, DRAM bytes, pipe utilisation, registers, ...)
  <out>_source_stalls.txt   warp-stall samples by CUDA source line (top lines)
usage: python tools/ncu_summarize.py gpurun_out/prof_sweep.ncu-rep profiles/r01_ncu_k_fast_sweep_window_regime
"""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = re.compile(r"gpu__time_duration|dram__bytes|dram__throughput|lts__t_bytes|lts__t_sectors_op|sm__throughput|"
                  r"sm__pipe_fp64|sm__inst_executed_pipe_fp64|smsp__inst_executed\.sum$|smsp__issue_active|"
                  r"launch__|sm__warps_active|smsp__sass_thread_inst_executed_op_d|l1tex__data_bank_conflicts|"
                  r"smsp__average_warp.*stall|smsp__average_warps_issue_stalled|sm__cycles_elapsed\.max|"
                  r"l1tex__data_pipe_lsu_wavefronts_mem_shared")


def ncu(*args):
    return subprocess.run(["ncu", "-i"] + list(args), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main(rep, out):
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    raw = {"Kernel Name": vals[hdr.index("Kernel Name")]}
    for h, u, v in zip(hdr, units, vals):
        if KEEP.search(h):
            raw[h] = v
            if u:
                raw[h + " [unit]"] = u
    with open(out + "_raw.json", "w") as fh:
        json.dump(raw, fh, indent=1)
    # cuda,sass view: one section per source file; rows that carry a line number are the per-line aggregates
    src = csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass")))
    lines, fname, h = [], "", None
    for r in src:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            h = r
            c_samp, c_inst = h.index("# Samples"), h.index("Instructions Executed")
            stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
            continue
        if h is None or not r[0].isdigit():
            continue
        try:
            samp = float(r[c_samp] or 0)
        except ValueError:
            continue
        if samp <= 0:
            continue
        st = sorted(((float(r[i]), c) for i, c in stall_cols if r[i] not in ("", "0", "-")), reverse=True)[:4]
        lines.append((samp, "%s:%s" % (fname, r[0]), r[c_inst], st, r[1].strip()[:100]))
    lines.sort(reverse=True)
    total = sum(x[0] for x in lines)
    with open(out + "_source_stalls.txt", "w") as fh:
        fh.write("warp-stall samples by source line (ncu --page source); total samples %d\n" % total)
        for samp, ln, inst, st, text in lines[:200]:
            fh.write("%9d (%4.1f%%) inst=%-12s %-60s | %s %s\n" % (
                samp, 100.0 * samp / max(total, 1), inst, " ".join("%s=%d" % (c[6:], v) for v, c in st), ln, text))
    print("wrote", out + "_raw.json", out + "_source_stalls.txt", "lines", len(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
