"""Condense an .ncu-rep (ncu --set full) into a small JSON: selected raw metrics + the hottest SASS lines."""
import csv, io, json, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_static',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum']
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
out = {'kernel': r[hdr.index('Kernel Name')]}
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        out[w] = (r[i] + ' ' + units[i]).strip()
# every warp-stall reason (per issue-active cycle), largest first
stalls = {h: float(r[i]) for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and r[i]}
out['stalls_per_issue'] = {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''): round(v, 3)
                           for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]}
for w in ('launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'launch__shared_mem_per_block_dynamic',
          'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.sum',
          'lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum', 'lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum', 'sm__cycles_active.avg'):
    if w in hdr:
        i = hdr.index(w)
        out[w] = (r[i] + ' ' + units[i]).strip()
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
if len(srows) > 2:
    h = srows[1]
    if 'Warp Stall Sampling (All Samples)' in h:
        k = h.index('Warp Stall Sampling (All Samples)'); s = h.index('Source')
        data = [x for x in srows[2:] if len(x) > k]
        tot = sum(float(x[k] or 0) for x in data) or 1.0
        out['hot_sass'] = [[x[s].strip(), round(float(x[k] or 0) / tot, 4)] for x in sorted(data, key=lambda x: -float(x[k] or 0))[:8]]
print(json.dumps(out, indent=1))
