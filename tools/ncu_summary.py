"""Summarise an .ncu-rep: headline metrics, opcode histogram and the hottest source lines (development helper)."""
import collections, csv, re, subprocess, sys

rep = sys.argv[1]
units_per_launch = float(sys.argv[2]) if len(sys.argv) > 2 else 60e6

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum', 'sm__cycles_elapsed.avg', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_not_selected_per_warp_active.pct',
        'smsp__warp_issue_stalled_wait_per_warp_active.pct', 'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct',
        'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct', 'smsp__warp_issue_stalled_membar_per_warp_active.pct',
        'smsp__warp_issue_stalled_drain_per_warp_active.pct', 'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct']
d = data[0]
print("kernel:", d[hdr.index('Kernel Name')][:90], d[hdr.index('Block Size')], d[hdr.index('Grid Size')])
for w in want:
    if w in hdr:
        print(f"  {w:75s} {d[hdr.index(w)]:>16s} {units[hdr.index(w)]}")

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
iex, ith = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
opc = collections.Counter(); tot = 0; tth = 0; n_static = 0
for r in rows[2:]:
    if len(r) <= iex:
        continue
    try:
        n = int(r[iex]); t = int(r[ith])
    except ValueError:
        continue
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[1])
    op = m.group(2).split('.')[0] if m else '?'
    opc[op] += n; tot += n; tth += t; n_static += 1
print(f"static SASS {n_static}, warp inst {tot}, thread inst/unit {tth / units_per_launch:.1f}, warp-inst per 32 units {tot * 32 / units_per_launch:.1f}")
for op, n in opc.most_common(26):
    print(f"  {op:10s} {n:>12} {100 * n / tot:5.1f}%  {n * 32 / units_per_launch:6.1f} per 32 units")

cs = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(cs.splitlines()))
hdr = rows[2]
iex = hdr.index('Instructions Executed')
lines = []
for r in rows[3:]:
    if len(r) > iex and r[0] != '':
        try:
            lines.append((int(r[iex]), int(r[0]), r[1].strip()[:120]))
        except ValueError:
            pass
lines.sort(reverse=True)
print("hottest source lines (inclusive of inlined callees, so lines overlap):")
for n, l, s in lines[:40]:
    print(f"  {n * 32 / units_per_launch:7.1f}  L{l}: {s}")
