#!/usr/bin/env python3
"""Summarise an ncu report (gpurun_out/*.ncu-rep) into profiles/<name>.md: the raw-page metrics the roofline uses and
the per-stage instruction mix from the source page.  Usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.md"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.avg.per_second", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    lines = ["# ncu summary of `%s`" % rep, ""]
    for row in raw[2:]:
        d = dict(zip(hdr, row)); u = dict(zip(hdr, units))
        lines.append("## kernel `%s`" % d.get("Kernel Name", "?"))
        lines.append("")
        lines.append("| metric | value | unit |"); lines.append("|---|---|---|")
        for k in KEYS:
            if k in d:
                lines.append("| %s | %s | %s |" % (k, d[k], u.get(k, "")))
        stall = {k.split("stalled_")[1]: int(v) for k, v in d.items() if "pcsamp_warps_issue_stalled" in k and not k.endswith("not_issued") and v.isdigit()}
        tot = sum(stall.values()) or 1
        lines.append(""); lines.append("warp-state samples: " + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:8]))
        try:
            rd = float(d["dram__bytes_read.sum"]); wr = float(d["dram__bytes_write.sum"])
            lines.append(""); lines.append("DRAM traffic per launch: read %s %s + write %s %s" % (d["dram__bytes_read.sum"], u["dram__bytes_read.sum"], d["dram__bytes_write.sum"], u["dram__bytes_write.sum"]))
        except Exception:
            pass
        lines.append("")
    src = page(rep, "source")
    if len(src) > 2:
        h = src[1]; ia = h.index("Instructions Executed"); isrc = h.index("Source"); isamp = h.index("# Samples")
        data = [r for r in src[2:] if len(r) > ia and r[ia].isdigit()]
        tot = sum(int(r[ia]) for r in data) or 1
        seg = collections.Counter(); ops = collections.defaultdict(collections.Counter); nb = 0
        for r in data:
            if "BAR.SYNC" in r[isrc]:
                nb += 1
            seg[nb] += int(r[ia])
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
            ops[nb][m.group(2).split(".")[0] if m else "?"] += int(r[ia])
        lines.append("## warp instructions by code region (split at BAR.SYNC), first kernel in the report")
        lines.append(""); lines.append("total %d warp instructions" % tot); lines.append("")
        lines.append("| region | warp instr | share | top opcodes |"); lines.append("|---|---|---|---|")
        for k in sorted(seg):
            lines.append("| %d | %d | %.1f%% | %s |" % (k, seg[k], 100 * seg[k] / tot, ", ".join("%s %d" % kv for kv in ops[k].most_common(8))))
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    main()
