"""Where the stage loop's instructions come from: SASS of a model's cubin attributed to source lines (nvdisasm --print-line-info),
restricted to the RK4 stage loop (the backward-branch loop with > 70 % FP64 instructions) and grouped by the region of
psy_filter.cuh / the model's equations they were compiled from.  No GPU needed.

    python tools/sass_by_source.py            # C4 model
"""
import collections
import re
import subprocess
import sys

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import psydiff_b200 as pd  # noqa: E402
from psydiff_b200 import synth  # noqa: E402

FILTER = open(__import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(pd.__file__)), "csrc", "psy_filter.cuh")).read().splitlines()


def _landmarks():
    """Ordered (first line, label) pairs for rhs_scaled and the register-resident rk4 loop of psy_filter.cuh."""
    marks = [("__device__ __forceinline__ double rcp(double x)", "reciprocals of diag(A): seed + Newton steps"),
             ("template <class Mdl>", None),
             ("static __device__ __forceinline__ void rhs_scaled(", "rhs_scaled prologue"),
             ("rd[i] = rcp(A[tri(i, i)]);", "reciprocals of diag(A): seed + Newton steps"),
             ("At[tri(i, j)] = A[tri(i, j)] * rd[j];", "column scaling: A-tilde = A D^-1"),
             ("double acc = -At[tri(i, j)];", "W = inverse of the unit-lower A-tilde"),
             ("Vec<N> xc, f0;", "drift at the centre point, seed of the mean"),
             ("Vec<N> xp, xm, fp, fm;", "sigma points m +/- sqrt(c) A_j"),
             ("Mdl::latentEquation(xp, fp, p, person, t);", "drift calls (inlined user equations attributed to the .cu file)"),
             ("double d[N];", "differences d_j = f(x+) - f(x-) and FMA accumulation of the mean"),
             ("double acc = d[i];", "G-tilde[:, j] = W d_j"),
             ("double Ph[TN], ka[N];", "Phi(M-tilde): diffusion term W Q W' + kap (G-tilde A_jj + A_ii G-tilde')"),
             ("// dA = (Ã Phi(M̃)) D⁻¹", "dA = A-tilde Phi(M-tilde)"),
             ("static __device__ __forceinline__ void rhs(", None),
             ("static __device__ __forceinline__ void rk4(", "RK4 loop control"),
             ("double mt[N], At[TN], ma[N], Aa[TN];", "RK4: stage input, running sum, step coefficients b D^-1, a D^-1"),
             ("// ---- adaptive Dormand-Prince", None)]
    out, pos = [], 0
    for text, label in marks:
        for k in range(pos, len(FILTER)):
            if text in FILTER[k]:
                out.append((k + 1, label))
                pos = k + 1
                break
    return out


LANDMARKS = _landmarks()


def region_of(path, line):
    if path.endswith(".cu"):
        return "user equations (drift statements of the model)"
    if path.endswith("psy_lang.cuh"):
        return "equation language (Vec / Mat operators)"
    if not path.endswith("psy_filter.cuh"):
        return path.rsplit("/", 1)[-1]
    label = "other (kernel prologue / time loop)"
    for first, lab in LANDMARKS:
        if line >= first:
            label = lab or "other (kernel prologue / time loop)"
    return label


def main():
    model = pd.newPsydiff(**synth.config_c4(N=8, T=4))
    cub = pd.compileModelSource(model)
    out = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
    cur = ("?", 0)
    ins = []                                              # (address, opcode text, (file, line))
    for l in out.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip(), cur))
    loops = []
    for a, t, _ in ins:
        if "BRA" in t:
            m2 = re.search(r"0x([0-9a-f]+)", t) or re.search(r"`\((\.L_x_\d+)\)", t)
            if m2 and m2.group(1).startswith(".L"):
                continue
            if m2 and int(m2.group(1), 16) < a:
                loops.append((int(m2.group(1), 16), a))
    # nvdisasm prints labels instead of addresses for branch targets: fall back to the densest FP64 window via execution structure
    op = lambda t: re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
    if not loops:
        # locate the stage loop as the longest run of instructions between two labels that is > 70 % FP64 and > 500 long
        text = out.splitlines()
        blocks, curb = [], []
        for l in text:
            if re.match(r"\s*\.L_x_\d+:", l):
                blocks.append(curb)
                curb = []
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m:
                curb.append(int(m.group(1), 16))
        blocks.append(curb)
        addr2ins = {a: (t, src) for a, t, src in ins}
        best = max((b for b in blocks if len(b) > 300), key=lambda b: sum(op(addr2ins[a][0]) in ("DFMA", "DMUL", "DADD") for a in b) / len(b))
        body = [(a,) + addr2ins[a] for a in best]
    else:
        lo, hi = min(loops, key=lambda p: p[1] - p[0])
        body = [x for x in ins if lo <= x[0] <= hi]
    by = collections.Counter()
    fp = collections.Counter()
    for a, t, (path, line) in body:
        r = region_of(path, line)
        by[r] += 1
        if op(t) in ("DFMA", "DMUL", "DADD"):
            fp[r] += 1
    tot = len(body)
    print(f"stage loop: {tot} instructions, {sum(fp.values())} FP64")
    for r, n in by.most_common():
        print(f"  {n:4d} ({100 * n / tot:4.1f} %)  FP64 {fp[r]:4d}   {r}")


if __name__ == "__main__":
    main()
