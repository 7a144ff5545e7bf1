"""FP64 instructions of a model kernel's RK4 stage loop (the shortest backward-branch loop that is > 70 % FP64) by source line.

    python tools/sass_loop_by_line.py c4|c2|c5|k<K> [extra flags]       (no GPU needed)
"""
import collections, os, re, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import psydiff_b200 as pd
from run_model import config

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
flags = " ".join(sys.argv[2:])
cub = pd.compileModelSource(pd.newPsydiff(**config(name, 8, 4)), flags)
out = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
ins, lab, cur = [], {}, None
for l in out.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s*(\.L_x_\d+):', l)
    if m:
        lab[m.group(1)] = len(ins)
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        ins.append((m.group(2).strip(), cur))
best = None
for i, (t, c) in enumerate(ins):
    m = re.search(r'BRA.*?(\.L_x_\d+)', t)
    if m and lab.get(m.group(1), i + 1) < i:
        body = ins[lab[m.group(1)]:i + 1]
        fp = sum(1 for t2, _ in body if re.match(r'(@!?U?P\d+\s+)?D(FMA|MUL|ADD)', t2))
        if fp > 0.7 * len(body) and len(body) > 100 and (best is None or len(body) < len(best)):
            best = body
srcs = {}
def text(f, ln):
    if f not in srcs:
        for d in (os.path.join(os.path.dirname(pd.__file__), "csrc"), os.path.dirname(cub)):
            if os.path.exists(os.path.join(d, f)):
                srcs[f] = open(os.path.join(d, f)).read().splitlines()
        srcs.setdefault(f, [])
    return srcs[f][ln - 1].strip()[:95] if 0 < ln <= len(srcs[f]) else ""
cnt = collections.defaultdict(collections.Counter)
for t, c in best:
    op = re.sub(r'^@!?U?P\d+\s+', '', t).split()[0].split('.')[0]
    cnt[c][op] += 1
tot = collections.Counter()
print(f"{name} {flags!r}: stage loop {len(best)} instructions")
for c in sorted(cnt, key=lambda x: (x or ("", 0))):
    k = cnt[c]
    tot.update(k)
    f, ln = c or ("?", 0)
    other = sum(v for o, v in k.items() if o not in ("DFMA", "DMUL", "DADD"))
    print(f"{k['DFMA']:4d} {k['DMUL']:4d} {k['DADD']:4d} {other:4d}  {f}:{ln}  {text(f, ln)}")
print("DFMA DMUL DADD other; totals:", tot["DFMA"], tot["DMUL"], tot["DADD"], sum(v for o, v in tot.items() if o not in ("DFMA", "DMUL", "DADD")))
