"""Top SASS lines by stall samples / shared wavefronts from `ncu --page source --csv --print-source sass` of one kernel.
usage: ncu -i rep --page source --csv --print-source sass > x.csv; python scripts/ncu_hot_sass.py x.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
# the file can hold several kernels: take the first block
start = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
end = next((i for i in range(start + 1, len(rows)) if rows[i] and rows[i][0] == 'Kernel Name'), len(rows))
hdr = rows[start]
body = [r for r in rows[start + 1:end] if len(r) >= len(hdr) - 2]
ci = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ci[k]].replace(',', ''))
    except Exception: return 0.0
tot = sum(f(r, '# Samples') for r in body)
wsh = sum(f(r, 'L1 Wavefronts Shared') for r in body)
print(f'total samples {tot:.0f}, shared wavefronts {wsh:.0f}, instr executed {sum(f(r, "Instructions Executed") for r in body):.0f}')
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(f(r, s) for r in body) for s in stalls}
print('stall mix:', ', '.join(f'{k[6:]}={v / max(tot, 1) * 100:.1f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
print('--- top by samples')
for r in sorted(body, key=lambda r: -f(r, '# Samples'))[:n]:
    top = max(stalls, key=lambda s: f(r, s))
    print(f'{f(r, "# Samples") / max(tot, 1) * 100:5.1f}%  {r[ci["Address"]][-5:]}  {r[ci["Source"]][:90]:90s} exec={f(r, "Instructions Executed"):.0f} wsh={f(r, "L1 Wavefronts Shared"):.0f} {top[6:]}')
print('--- top by shared wavefronts')
for r in sorted(body, key=lambda r: -f(r, 'L1 Wavefronts Shared'))[:10]:
    print(f'{f(r, "L1 Wavefronts Shared") / max(wsh, 1) * 100:5.1f}%  {r[ci["Source"]][:90]:90s} exec={f(r, "Instructions Executed"):.0f} wsh={f(r, "L1 Wavefronts Shared"):.0f} ideal={f(r, "L1 Wavefronts Shared Ideal"):.0f}')
