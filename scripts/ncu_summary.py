"""Condense an `ncu --set full` report into the handful of numbers the roofline discussion needs.
usage: python scripts/ncu_summary.py report.ncu-rep > profiles/rNN_<what>.txt"""
import csv, io, subprocess, sys
KEYS = [('gpu__time_duration.sum', 'duration'),
        ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram % of peak'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 % of peak'),
        ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1/tex % of peak'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe % (active)'),
        ('sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active', 'legacy HMMA pipe %'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
        ('launch__registers_per_thread', 'registers/thread'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('launch__shared_mem_per_block_dynamic', 'dyn smem/block')]
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('kernel:', r[hdr.index('Kernel Name')][:150])
    for k, label in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f'    {label:28s} {r[i]} {units[i]}')
    if 'dram__bytes_read.sum' in hdr:
        def val(k):
            i = hdr.index(k); v = float(r[i].replace(',', '')); u = units[i]
            return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        print(f'    {"dram traffic (r+w)":28s} {(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / 1e6:.1f} MB')
    print()
