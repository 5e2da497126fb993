#!/usr/bin/env python
"""Static look at a kernel's hottest loop in SASS: instruction mix and, for every instruction, the distance
(in instructions) to the producer of its operands. Short distances between dependent packed FP32 ops are
what shows up as stall_wait in ncu.   usage: sass_deps.py <object-or-so> <mangled-name-substring>"""
import re, subprocess, sys, collections
obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
f = [x for x in funcs if pat in x.split("\n")[0]][0]
ins = []
for l in f.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
# hottest loop = backward branch with the largest body containing FFMA2
best = None
for k, (a, t) in enumerate(ins):
    m = re.search(r"BRA\s+(?:\w+,\s*)?0x([0-9a-f]+)", t)
    if m and int(m.group(1), 16) < a:
        tgt = int(m.group(1), 16)
        body = [(x, y) for x, y in ins if tgt <= x <= a]
        n = sum(1 for _, y in body if "FFMA2" in y)
        if best is None or n > best[0]: best = (n, body)
body = best[1]
print("loop body: %d instructions" % len(body))
mix = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in body)
print(" ".join("%s=%d" % kv for kv in mix.most_common()))
def regs(tok):
    out = []
    for m in re.finditer(r"\bR(\d+)(\.F32x2|\.64)?", tok):
        r = int(m.group(1)); out.append(r)
        if m.group(2): out.append(r + 1)
    return out
last = {}
dist = collections.Counter()
n2 = len(body)
for rep in range(2):           # two passes so loop-carried producers are seen
    for k, (a, t) in enumerate(body):
        t2 = re.sub(r"^@!?U?P\d+\s+", "", t)
        op = t2.split()[0]
        ops = t2[len(op):].split(",")
        dst = ops[0]; srcs = ",".join(ops[1:])
        wide = op.startswith(("FADD2", "FMUL2", "FFMA2"))
        pos = rep * n2 + k
        if rep == 1 and op.split(".")[0] in ("FADD2", "FMUL2", "FFMA2", "FMNMX3"):
            sr = []
            for m in re.finditer(r"\bR(\d+)(\.F32x2\.HI_LO|\.F32)?", srcs):
                r = int(m.group(1)); sr.append(r)
                if m.group(2) == ".F32x2.HI_LO" or (op.startswith("FMNMX3") is False and wide and m.group(2) is None): sr.append(r + 1)
            d = min([pos - last[r] for r in sr if r in last] or [99])
            dist[(op.split(".")[0], min(d, 8))] += 1
        for m in re.finditer(r"\bR(\d+)", dst):
            r = int(m.group(1)); last[r] = pos
            if wide or "LDS.128" in op or ".64" in op:
                last[r + 1] = pos
            if "LDS.128" in op: last[r + 2] = pos; last[r + 3] = pos
for op in ("FADD2", "FMUL2", "FFMA2", "FMNMX3"):
    print(op, " ".join("d%d:%d" % (d, dist[(op, d)]) for d in range(1, 9) if dist[(op, d)]))
