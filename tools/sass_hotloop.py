"""Static proxy for kernel changes without a GPU: instruction mix of the RK4 stage loop (the hot loop) of a model's cubin."""
import collections, re, subprocess, sys
sys.path.insert(0, '/root/repo')
import psydiff_b200 as pd
from psydiff_b200 import synth


def hotloop(cubin):
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    ins = []
    for l in sass.splitlines():
        m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    loops = []
    for a, t in ins:
        if 'BRA' in t:
            m2 = re.search(r'0x([0-9a-f]+)', t)
            if m2 and int(m2.group(1), 16) < a:
                loops.append((int(m2.group(1), 16), a))
    best = None
    for lo, hi in loops:
        body = [t for a, t in ins if lo <= a <= hi]
        c = collections.Counter(re.sub(r'^@!?U?P\d+\s+', '', t).split()[0].split('.')[0] for t in body)
        fp = c['DFMA'] + c['DMUL'] + c['DADD']
        if fp > 0.7 * len(body) and (best is None or len(body) < best[0]) and len(body) > 200:
            best = (len(body), fp, c)
    return best, len(ins)


if __name__ == "__main__":
    flags = sys.argv[1] if len(sys.argv) > 1 else ""
    for name, kw in (("C4", synth.config_c4(N=8, T=4)), ("n8", synth.config_coupled(K=4, N=8, T=4))):
        m = pd.newPsydiff(**kw)
        cub = pd.compileModelSource(m, flags)
        log = open(cub.replace(".cubin", ".log")).read()
        spill = re.search(r"(\d+) bytes stack frame", log).group(1)
        (n, fp, c), total = hotloop(cub)
        print(f"{name}: hot loop {n} instrs, FP64 {fp} (DFMA {c['DFMA']} DMUL {c['DMUL']} DADD {c['DADD']}), LDL {c['LDL']} STL {c['STL']} MUFU {c['MUFU']}; stack {spill} B; kernel {total} instrs")
