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
        roles = any("PMTRIG" in r[isrc] for r in data)       # warp-specialised kernel: split at the role markers (pmevent), else at BAR.SYNC
        seg = collections.Counter(); ops = collections.defaultdict(collections.Counter); smp = collections.Counter(); poll = collections.Counter(); nb = "prologue" if roles else 0
        stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
        stalls = collections.defaultdict(collections.Counter)
        ROLE = {"0x2": "role P (raw codes -> floats, chroma reconstruction; issues the TMA loads)", "0x4": "role A (convert shader body)",
                "0x8": "role B (X pass)", "0x10": "role C (Y pass, rounding, dither, stores)"}
        for r in data:
            if roles:
                m = re.search(r"PMTRIG (0x[0-9a-f]+)", r[isrc])
                if m: nb = ROLE.get(m.group(1), m.group(1))
            elif "BAR.SYNC" in r[isrc]:
                nb += 1
            n = int(r[ia]); seg[nb] += n; smp[nb] += int(r[isamp] or 0)
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
            op = m.group(2).split(".")[0] if m else "?"
            ops[nb][op] += n
            if op in ("SYNCS", "NANOSLEEP"): poll[nb] += n
            for i, name in stall_cols:
                if i < len(r) and r[i].isdigit(): stalls[nb][name.replace("stall_", "")] += int(r[i])
        lines.append("## warp instructions by code region (%s), first kernel in the report" % ("split at the pmevent role markers" if roles else "split at BAR.SYNC"))
        lines.append(""); lines.append("total %d warp instructions%s" % (tot, " (source-page counters; the replay pass that collects them runs slower, so spin-wait"
                                       " instructions are over-represented against `smsp__inst_executed.sum` above)" if roles else "")); lines.append("")
        lines.append("| region | warp instr | share | of which mbarrier polling (SYNCS + NANOSLEEP) | samples | top stall reasons (samples) | top opcodes |"); lines.append("|---|---|---|---|---|---|---|")
        for k in seg:
            st = stalls[k]; stt = sum(st.values()) or 1
            lines.append("| %s | %d | %.1f%% | %d | %d | %s | %s |" % (k, seg[k], 100 * seg[k] / tot, poll[k], smp[k],
                         ", ".join("%s %.0f%%" % (a, 100 * b / stt) for a, b in st.most_common(5)), ", ".join("%s %d" % kv for kv in ops[k].most_common(8))))
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    main()
