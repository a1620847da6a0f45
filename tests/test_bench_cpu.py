"""bench.py contract checks that need no GPU: the reference (CPU) arm prints exactly one JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--config', 'tiny', '--steps', '2',
                        '--warmup', '1'], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j['impl'] == 'reference' and j['metric'] == 'tokens_per_sec' and j['unit'] == 'tokens/s' and j['higher_is_better'] is True
    assert j['value'] > 0 and j['cpu_baseline']['kind'] == 'port' and j['cpu_baseline']['cores'] >= 1
    assert j['e2e']['h2d_bytes_per_step'] == 0 and j['e2e']['d2h_bytes_per_step'] == 0 and j['e2e']['value'] == j['value']


def test_flop_model_matches_survey_table():
    sys.path.insert(0, ROOT)
    import bench
    cfg2 = bench.CONFIGS['cfg2']['kwargs']
    assert abs(bench.fwd_flops_per_token(cfg2) / 1e6 - 110.38) < 0.01          # cfg2 F_fwd = 110.38 MFLOP/token
    cfg3 = bench.CONFIGS['cfg3']['kwargs']
    assert abs(bench.fwd_flops_per_token(cfg3) / 1e6 - 881.38) < 0.05          # cfg3: 881.38
