"""pretty-print the JSON lines of scripts/decode_profile.py"""
import json, sys
for line in open(sys.argv[1]):
    o = json.loads(line)
    print('B', o['batch'], 'tok/s', o['tokens_per_sec'], 'us/step', o['us_per_step'])
    for k in ('pos300', 'pos900'):
        print(' ', k, 'step cycles', o[k]['step_cycles'])
        for nm, r in o[k]['per_phase_avg_cycles'].items():
            print('    %-8s n=%2d  cta0 comp %6d wait %6d | ctaN comp %6d wait %6d' % (nm, r['n'], r['cta0_compute'], r['cta0_wait'], r['ctaN_compute'], r['ctaN_wait']))
        for nm, m in o[k].get('marks_enter_staged_fma_final_prefetch_ln1_ln2', {}).items():
            print('    marks %-8s enter %5d staged %5d fma %5d final %5d prefetch %5d | m5 %5d m6 %5d m7 %5d' % (nm, m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7]))
