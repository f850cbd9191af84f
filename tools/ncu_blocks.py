"""Development helper: basic-block view of an ncu SASS source page -- runs of instructions with the same execution count,
with their share of executed warp instructions and of stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]) if len(sys.argv) > 2 else 60e6
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.4
hdr = rows[1]
ia, isrc, iex, ith, ismp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
ins = []
for r in rows[2:]:
    try:
        ins.append((int(r[ia], 16), r[isrc].strip(), int(r[iex]), int(r[ith]), int(r[ismp])))
    except (ValueError, IndexError):
        pass
base = ins[0][0]
tot = sum(i[2] for i in ins); smp = sum(i[4] for i in ins)
runs = []
for i in ins:
    if runs and runs[-1]['n'] == i[2]:
        r = runs[-1]; r['end'] = i[0]; r['k'] += 1; r['ex'] += i[2]; r['th'] += i[3]; r['s'] += i[4]
    else:
        runs.append({'start': i[0], 'end': i[0], 'n': i[2], 'k': 1, 'ex': i[2], 'th': i[3], 's': i[4], 'first': i[1]})
print(f"total warp inst {tot}  ({tot * 32 / units:.1f} per 32 units), samples {smp}")
for r in runs:
    share = 100 * r['ex'] / tot
    if share >= min_share or 100 * r['s'] / smp >= min_share:
        print(f"  {r['start'] - base:#7x}..{r['end'] - base:#7x}  static {r['k']:4d}  exec/instr {r['n']:>10}  inst {share:5.1f}% ({r['ex'] * 32 / units:6.1f}/32u)  samples {100 * r['s'] / smp:5.1f}%  thr/inst {r['th'] / max(r['ex'], 1):4.1f}  {r['first'][:60]}")
