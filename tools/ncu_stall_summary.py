"""Stall-reason / opcode summary of a kernel from an `ncu --set full --import-source on` report (no GPU needed).

    python tools/ncu_stall_summary.py gpurun_out/prof.ncu-rep

Reads the source page (`ncu -i REPORT --page source --csv --print-source sass`), splits the SASS by execution count into the RK4
stage loop (executed most), the per-step and the per-time-point code, and prints samples per stall reason and per opcode for
each.  This is how profiles/r01d_psy_filter_kernel_c4_stall_summary.txt was made.
"""
import collections
import csv
import io
import re
import subprocess
import sys


def load(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr, data = rows[start], rows[start + 1:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    per = []
    for r in data:
        if len(r) < len(hdr) or not r[ix["Address"]].startswith("0x"):
            continue
        per.append({"src": r[ix["Source"]].strip(), "samples": int(r[ix["# Samples"]] or 0), "exec": int(r[ix["Instructions Executed"]] or 0),
                    "stalls": {h: int(r[ix[h]] or 0) for h in stalls}})
    return per


def opcode(src):
    return re.sub(r"^@!?U?P\d+\s+", "", src).split()[0].split(".")[0]


def report(name, part, total):
    n = sum(p["samples"] for p in part)
    print(f"\n== {name}: {len(part)} instructions, {n} samples = {100 * n / total:.1f} % of the kernel")
    st = collections.Counter()
    for p in part:
        st.update(p["stalls"])
    for h, v in st.most_common(8):
        print(f"   {h:26s} {100 * v / max(n, 1):5.1f} %")
    by, cnt = collections.Counter(), collections.Counter()
    for p in part:
        by[opcode(p["src"])] += p["samples"]
        cnt[opcode(p["src"])] += 1
    print("   opcode      instrs   samples %")
    for op, v in by.most_common(10):
        print(f"   {op:10s} {cnt[op]:6d}   {100 * v / max(n, 1):6.2f}")


def main():
    per = load(sys.argv[1])
    total = sum(p["samples"] for p in per)
    mx = max(p["exec"] for p in per)
    report("whole kernel", per, total)
    report("stage loop (executed most)", [p for p in per if p["exec"] > 0.5 * mx], total)
    report("per RK4 step", [p for p in per if 0.1 * mx < p["exec"] <= 0.5 * mx], total)
    report("per time point and rarer", [p for p in per if 0 < p["exec"] <= 0.1 * mx], total)


if __name__ == "__main__":
    main()
