"""Static view of a cubin without a GPU: every backward-branch loop of psy_filter_kernel with its instruction mix (innermost = the
RK4 stage loop).  usage: python tools/sass_loops.py CUBIN"""
import collections, re, subprocess, sys


def loops(cubin):
    sass = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True).stdout
    ins = []
    for l in sass.splitlines():
        m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    out = []
    for a, t in ins:
        if 'BRA' in t:
            m2 = re.search(r'0x([0-9a-f]+)', t)
            if m2 and int(m2.group(1), 16) < a:
                lo = int(m2.group(1), 16)
                body = [x for b, x in ins if lo <= b <= a]
                c = collections.Counter(re.sub(r'^@!?U?P\d+\s+', '', x).split()[0].split('.')[0] for x in body)
                out.append((lo, a, len(body), c))
    return out, len(ins)


if __name__ == "__main__":
    ls, total = loops(sys.argv[1])
    print(f"kernel: {total} instructions")
    for lo, hi, n, c in sorted(ls, key=lambda x: x[2]):
        if n < 100:
            continue
        fp = c['DFMA'] + c['DMUL'] + c['DADD']
        rest = {k: v for k, v in c.most_common() if k not in ('DFMA', 'DMUL', 'DADD')}
        print(f"loop {lo:#x}-{hi:#x}: {n} instrs, FP64 {fp} (DFMA {c['DFMA']} DMUL {c['DMUL']} DADD {c['DADD']}); " + " ".join(f"{k} {v}" for k, v in list(rest.items())[:14]))
