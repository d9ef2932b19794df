#!/usr/bin/env python3
"""tools/make_capture_json.py <report.ncu-rep> [kernel-regex] -> profiles/capture_c2.json
Records, from one `ncu --set full` capture of the C2 frame kernel, the numbers bench.py quotes next to its live timing
(DRAM bytes per launch, warp instructions per launch) TOGETHER WITH the sha256 of the device sources (videorenderer_b200/csrc/*.cu, *.cuh, *.h) at capture time.
bench.py recomputes that hash: if the kernel sources changed since the capture it says so ("traffic_stale": true) and drops
the instruction-issue figure instead of quoting numbers of another binary."""
import csv, hashlib, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def csrc_sha256():
    d = os.path.join(ROOT, "videorenderer_b200", "csrc")
    h = hashlib.sha256()
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh", ".h")):            # device code and the headers it includes; host-only .cpp files do not change a kernel
            h.update(name.encode()); h.update(open(os.path.join(d, name), "rb").read())
    return h.hexdigest()


def kernel_sass_sha256(kernel_regex=r"k_strip2xINS_8StripCfgILi128ELi8ELi8ELi12ELi4ELi8ELi19EEELi6ELb1ELi1ELi0E"):
    """sha256 over the SASS instruction text (cuobjdump -sass, addresses and encodings stripped) of the captured frame kernel as it sits in
    the built library: the identity of the BINARY the capture's counters belong to.  An edit to a header that the kernel does not
    instantiate leaves it unchanged; None when cuobjdump or the library is missing."""
    lib = os.path.join(ROOT, "videorenderer_b200", "libmpcvr_b200.so")
    try:
        sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, timeout=300).stdout
    except Exception:
        return None
    h = hashlib.sha256(); cur = False; n = 0
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = bool(re.search(kernel_regex, m.group(1))); continue
        if cur:
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?)\s*/\* 0x[0-9a-f]+ \*/", ln)
            if m:
                h.update(m.group(1).encode()); h.update(b"\n"); n += 1
    return h.hexdigest() if n else None


def main():
    rep = sys.argv[1]; pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else "k_strip2x|k_fused2x")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for row in rows[2:]:
        d = dict(zip(hdr, row)); u = dict(zip(hdr, units))
        if not pat.search(d.get("Kernel Name", "")):
            continue
        def b(k):
            v = float(d[k]); unit = u.get(k, "")
            return int(round(v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(unit, 1)))
        cap = {"kernel": d["Kernel Name"], "report": os.path.basename(rep), "dram_bytes_read": b("dram__bytes_read.sum"), "dram_bytes_write": b("dram__bytes_write.sum"),
               "warp_instr_per_launch": int(float(d["smsp__inst_executed.sum"])), "gpu_time_us": float(d["gpu__time_duration.sum"]),
               "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]), "registers_per_thread": int(float(d["launch__registers_per_thread"])),
               "csrc_sha256": csrc_sha256(), "kernel_sass_sha256": kernel_sass_sha256()}
        cap["dram_bytes_per_launch"] = cap["dram_bytes_read"] + cap["dram_bytes_write"]
        # packed fp32 instructions (two FMA-pipe cycles each) from the source page of the same report: the input of the issue-cost model
        # in DESIGN.md 4.1 (about 2.35 scheduler cycles per packed, 1.1 per other warp instruction; tools/role_bench.cu)
        try:
            src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)))
            hi = [i for i, r in enumerate(src) if "Instructions Executed" in r][0]
            ia, isrc = src[hi].index("Instructions Executed"), src[hi].index("Source")
            cap["packed_warp_instr_per_launch"] = sum(int(r[ia]) for r in src[hi + 1:] if len(r) > ia and r[ia].isdigit() and re.search(r"\b(FFMA2|FADD2|FMUL2)\b", r[isrc]))
        except Exception:
            pass
        json.dump(cap, open(os.path.join(ROOT, "profiles", "capture_c2.json"), "w"), indent=1)
        print(json.dumps(cap, indent=1)); return 0
    print("no kernel matching", pat.pattern); return 1


if __name__ == "__main__":
    sys.exit(main())
