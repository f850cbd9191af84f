"""Static view of a kernel's SASS loops (development helper): python tools/sass_loops.py <lib.so> <kernel substring> [max body size] [print]
Lists every loop (backward branch) smallest first with its opcode mix; `print` dumps the bodies too."""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
limit = int(sys.argv[3]) if len(sys.argv) > 3 else 400
show = len(sys.argv) > 4
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
for b in txt.split("Function : ")[1:]:
    name = b.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for line in b.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print(name[:110], "static instructions:", len(ins))
    idx = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        if "BRA" in t:
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in idx:
                loops.append((idx[int(m.group(1), 16)], i))
    for s, e in sorted(loops, key=lambda x: x[1] - x[0]):
        body = ins[s:e + 1]
        if len(body) < 20 or len(body) > limit:
            continue
        ops = collections.Counter(re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", t).group(2).split(".")[0] for _, t in body)
        print(f"  loop {ins[s][0]:#x}..{ins[e][0]:#x}: {len(body)} instr  ", ops.most_common(16))
        if show:
            for a, t in body:
                print(f"      {a:#06x}  {t}")
    break
