"""The JSON line bench.py prints is a contract with the driver: the stored lines of the last measured runs
(profiles/r1_runs/) must carry every key the contract names, with sane values. CPU-only."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNS = os.path.join(ROOT, 'profiles', 'r1_runs')


def _line(name):
    path = os.path.join(RUNS, name)
    if not os.path.exists(path):
        pytest.skip('%s not recorded' % name)
    lines = [l for l in open(path).read().strip().splitlines() if l.startswith('{')]
    assert len(lines) == 1, 'exactly one JSON line per run'
    return json.loads(lines[0])


@pytest.mark.parametrize('name,n', [('bench_n1.json', 1), ('bench_n2.json', 2), ('bench_n4.json', 4), ('bench_n8.json', 8)])
def test_bench_line_schema(name, n):
    d = _line(name)
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'clocks', 'roofline'):
        assert k in d, k
    assert d['metric'] == 'agent_steps_per_sec' and d['unit'] == 'agent-steps/s' and d['n_gpus'] == n
    assert d['higher_is_better'] is True and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert d['dtype'] == 'f32' and d['data'] == 'synthetic' and 'workload' in d['config'] and d['warmup'] >= 3
    assert d['gpu_launches'] > 0
    e = d['e2e']
    assert e['unit'] == d['unit'] and e['value'] > 0 and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0
    r = d['roofline']
    assert r['peak'] > 0 and abs(r['frac'] - r['achieved'] / r['peak']) < 1e-9
    c = d['clocks']
    assert c['sm_mhz'] > 0.9 * c['sm_max_mhz'], 'clocks were not held back during the timed region'
    assert not any('slowdown' in x for x in c['reasons'])
    if n == 1:
        b = d['cpu_baseline']
        assert b['kind'] in ('reference', 'port') and b['cores'] >= 1 and b['value'] > 0 and b['sample']


def test_replicate_ensemble_scales_linearly():
    one, eight = _line('bench_n1.json'), _line('bench_n8.json')
    assert eight['value'] / one['value'] > 7.2     # replicates are independent: no collective on the data path
