"""Static view of a kernel's SASS loops (development helper): python tools/sass_loops.py <lib.so> <kernel substring>"""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = txt.split("Function : ")
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for line in b.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print(name[:110], "static instructions:", len(ins))
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?`?\(?\.?L?_?x?_?([0-9a-f]+)?\)?|BRA.*?0x([0-9a-f]+)", t)
        m2 = re.search(r"0x([0-9a-f]+)", t) if "BRA" in t else None
        if m2:
            target = int(m2.group(1), 16)
            if target <= a and target in addr_index:
                loops.append((addr_index[target], i))
    for s, e in sorted(loops, key=lambda x: -(x[1] - x[0]))[:8]:
        body = ins[s:e + 1]
        ops = collections.Counter(re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", t).group(2).split(".")[0] for _, t in body)
        print(f"  loop {ins[s][0]:#x}..{ins[e][0]:#x}: {len(body)} instr  ", ops.most_common(14))
    break
