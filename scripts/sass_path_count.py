#!/usr/bin/env python
"""
Instruction counts of the far-field sweep's per-piece loop from the SASS of a build (no GPU needed).

    python scripts/sass_path_count.py tesseroid_variable_density_b200/csrc/_obj/tvd_far.o _Z9far_sweepILi8ELi512ELi1ELb0ELb1E [list]

Finds the loop of the tile-level far path (8 MUFU.RSQ64H, at most the one vote of the trig code) in
the given instantiation and walks the path of a further radial piece of the same tesseroid
(conditional forward branches that skip the longitude / latitude blocks are taken), printing the
counts by opcode.  Used for profiles/far_experiments_r2.md.
"""
import re, subprocess, sys

obj, func = sys.argv[1], sys.argv[2]
show = len(sys.argv) > 3
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
blocks = txt.split("Function : ")
body = [b for b in blocks if b.startswith(func) or func in b.split("\n")[0]]
assert body, "function not found"
body = body[0]
ins = []
for line in body.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr2i = {a: i for i, (a, _) in enumerate(ins)}


def cls(op):
    op = re.sub(r"^@!?U?P\d+\s+", "", op)
    mn = op.split()[0].split(".")[0]
    if mn in ("DADD", "DMUL", "DFMA", "DSETP"):
        return "fp64"
    return mn


loops = []
for i, (a, op) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", op)
    if m and "BRA" in op:
        t = int(m.group(1), 16)
        if t <= a and t in addr2i:
            j = addr2i[t]
            ops = [o for _, o in ins[j:i + 1]]
            nm = sum("MUFU.RSQ64H" in o for o in ops)
            nv = sum("VOTE" in o for o in ops)
            loops.append((t, a, nm, nv, len(ops)))
print("loops (start, end, mufu, vote, n):", [(hex(s), hex(e), m, v, n) for s, e, m, v, n in loops])
cand = [l for l in loops if l[2] >= 8 and l[3] <= 1]
assert cand, "tile_far loop not found"
s, e, _, _, _ = cand[0]
# walk the per-piece path
i = addr2i[s]
path = []
guard = 0
while True:
    a, op = ins[i]
    path.append((a, op))
    guard += 1
    if guard > 2000:
        break
    if a == e:
        break
    m = re.search(r"^(@!?U?P\d+\s+)?BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", op)
    if m:
        t = int(m.group(2), 16)
        cond = m.group(1) is not None or "UP" in op
        if t > a and t <= e + 0x10:
            # forward branch inside loop: taken if the skipped code has no MUFU.RSQ64H and is non-empty
            skipped = [o for aa, o in ins[i + 1:addr2i[t]]]
            if not cond or not any("MUFU.RSQ64H" in o for o in skipped):
                i = addr2i[t]
                continue
        elif t > e:      # loop exit: not taken
            pass
        elif t <= a:     # backward inside loop (not the end)
            pass
    i += 1
cnt = {}
for a, op in path:
    c = cls(op)
    cnt[c] = cnt.get(c, 0) + 1
nfp = cnt.get("fp64", 0)
noth = len(path) - nfp
print("per-piece path: %d instr: fp64 %d, other %d -> issue cycles 2*fp64+other = %d" % (len(path), nfp, noth, 2 * nfp + noth))
print(sorted(cnt.items(), key=lambda kv: -kv[1]))
if show:
    for a, op in path:
        print("%04x %s" % (a, op))
