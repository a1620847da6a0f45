"""Summarise an `ncu --set full` report: one block per distinct kernel (median over its launches) with duration, DRAM bytes,
pipe utilisation, occupancy and the top warp-stall reasons.
usage: python scripts/ncu_summary.py REPORT.ncu-rep [OUT.txt]"""
import csv, io, statistics, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {n: i for i, n in enumerate(hdr)}
def num(r, k):
    try:
        return float(r[ix[k]].replace(',', ''))
    except Exception:
        return None
groups = {}
for r in data:
    groups.setdefault(r[ix['Kernel Name']], []).append(r)
METRICS = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput % of peak'),
    ('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'tensor (hmma) pipe active %'),
    ('sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active', 'tensor pipe inst %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe cycles active %'),
    ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'XU (MUFU) pipe %'),
    ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'FMA pipe %'),
    ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'ALU pipe %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('launch__registers_per_thread', 'registers / thread'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic smem / block'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate %'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'shared bank conflicts'),
]
out = [f'source: {rep}  (ncu --set full --clock-control none; median over the launches of each kernel)', '']
for name, rs in groups.items():
    out.append(f'== {name}')
    out.append(f'   launches captured: {len(rs)}')
    for k, label in METRICS:
        if k in ix:
            vals = [v for v in (num(r, k) for r in rs) if v is not None]
            if vals:
                out.append(f'   {label:34s} {statistics.median(vals):14.3f} {units[ix[k]]}')
    stalls = []
    for k in ix:
        if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio') or \
           (k.startswith('smsp__average_warp_latency_issue_stalled_') and k.endswith('.ratio')):
            vals = [v for v in (num(r, k) for r in rs) if v is not None]
            if vals:
                stalls.append((statistics.median(vals), k.replace('smsp__average_warps_issue_stalled_', '').replace('smsp__average_warp_latency_issue_stalled_', '').replace('_per_issue_active.ratio', '').replace('.ratio', '')))
    stalls.sort(reverse=True)
    if stalls:
        out.append('   top warp stalls (warps per issue): ' + ', '.join(f'{n} {v:.2f}' for v, n in stalls[:6]))
    out.append('')
txt = '\n'.join(out)
if len(sys.argv) > 2:
    open(sys.argv[2], 'w').write(txt + '\n')
else:
    print(txt)
