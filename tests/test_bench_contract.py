"""The bench line contract, checked on the lines committed under profiles/ (CPU only: nothing is measured here).
A change of bench.py that drops or renames a key the driver reads shows up as a failure once new lines are committed."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line(name):
    with open(os.path.join(ROOT, 'profiles', name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


@pytest.mark.parametrize('name', ['r1_bench_c1_v28.json', 'r1_bench_c2_v27_run1.json', 'r1_bench_c2_v27_run2.json'])
def test_b200_line_has_every_contract_key(name):
    d = line(name)
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'roofline', 'cpu_baseline', 'e2e', 'gpu_launches', 'clocks'):
        assert k in d, k
    assert d['warmup'] >= 3 and d['steps'] >= 1 and d['higher_is_better'] is True and d['data'] == 'synthetic'
    assert d['vs_baseline'] is None                       # BASELINE.md publishes no number for this metric
    assert 'workload' in d['config'] and 'model' not in d['config']
    r = d['roofline']
    assert r['bound'] in ('hbm', 'tensor') and r['unit'] in ('GB/s', 'TFLOP/s')
    assert r['frac'] == pytest.approx(r['achieved'] / r['peak'])
    assert r['achieved'] == pytest.approx(r['algorithmic_bytes_per_launch'] / (r['ms_per_launch'] / 1e3) / 1e9)
    c = d['cpu_baseline']
    assert c['kind'] in ('reference', 'port') and c['cores'] >= 1 and c['value'] > 0 and c['sample']
    e = d['e2e']
    assert e['value'] > 0 and e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] > 0
    assert e['value'] != d['value']                       # the end-to-end arm is its own measurement
    assert d['gpu_launches'] > 0
    assert d['clocks']['sm_mhz'] and not set(d['clocks']['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}


def test_default_line_carries_config2_from_a_fresh_process():
    c = line('r1_bench_c1_v28.json')['config2_coverage_stage']
    assert 'error' not in c and c['workload'].startswith('C2')
    assert c['ms_per_step'] > 0 and c['e2e']['h2d_bytes_per_step'] == 120_000_000
    assert c['roofline']['kernel'] == 'cov_bin_events' and c['roofline']['traffic'] > 0


def test_reference_line():
    d = line('r1_bench_reference_v25.json')
    assert d['impl'] == 'reference' and d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
