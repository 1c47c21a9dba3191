"""bench.py contract checks that need no GPU: the reference arm (the reference's algorithm on the host cores) prints one
JSON line with the keys the driver reads, on the same metric / unit as the engine arm; defaults finish in minutes."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                         check=True, capture_output=True, text=True, cwd=ROOT, timeout=600).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'terms/s' and d['higher_is_better'] is True
    assert d['n_gpus'] == 1 and d['steps'] == 1 and d['dtype'] == 'f64' and d['data'] == 'synthetic'
    assert d['value'] > 1.0e5 and d['gpu_launches'] == 0 and d['vs_baseline'] is None
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {"value": d['value'], "unit": "terms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    base = json.load(open(os.path.join(ROOT, 'BASELINE.json')))
    assert 'terms' in d['metric'] and 'terms' in json.dumps(base.get('metric', base))


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
                        '--warmup', '0'], capture_output=True, text=True, cwd=ROOT, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_defaults():
    sys.path.insert(0, ROOT)
    import bench
    old = sys.argv
    try:
        sys.argv = ['bench.py']
        a = bench.parse()
    finally:
        sys.argv = old
    assert a.gpus == 1 and a.steps >= 1 and a.warmup >= 3 and a.impl == 'engine' and a.kind == 'free' and a.precision == 'f64'


def test_committed_engine_line_carries_every_contract_key():
    """The last engine-arm line measured on a B200 (profiles/) has the keys the driver and the judge read."""
    d = json.load(open(os.path.join(ROOT, 'profiles', 'r01_bench_1e7x1024_final.json')))
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'clocks', 'roofline', 'cpu_baseline'):
        assert k in d, k
    assert d['config']['workload'] and d['warmup'] >= 3 and d['gpu_launches'] > 0 and d['dtype'] == 'f64'
    assert set(('value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step')) <= set(d['e2e'])
    assert d['e2e']['h2d_bytes_per_step'] > 0 and d['e2e']['d2h_bytes_per_step'] > 0 and d['e2e']['value'] != d['value']
    assert set(('sm_mhz', 'sm_max_mhz', 'reasons')) <= set(d['clocks'])
    r = d['roofline']
    assert set(('bound', 'achieved', 'peak', 'unit', 'frac', 'traffic')) <= set(r)
    assert abs(r['frac'] - r['achieved'] / r['peak']) < 1e-12 and 0.5 < r['frac'] < 1.0
    assert set(('value', 'unit', 'cores', 'kind', 'sample')) <= set(d['cpu_baseline'])


def test_committed_round2_lines_carry_strong_scaling_and_parity():
    """Round 2: the default line is the strong-scaling run of the target catalogue and every line (1, 2, 8 GPUs) carries the
    parity block measured in that run, plus the sub-results the multi-GPU lines add."""
    for name, n in (('r02_bench_1gpu_final.json', 1), ('r02_bench_2gpu.json', 2), ('r02_bench_8gpu.json', 8)):
        d = json.loads(open(os.path.join(ROOT, 'profiles', name)).read().strip().splitlines()[-1])
        assert d['n_gpus'] == n and d['scaling'] == 'strong' and d['config']['sources_total'] == 10000000
        assert d['config']['sources_per_gpu'] * n == 10000000 and d['config']['walkers'] == 1024
        p = d['parity']
        assert p['ok'] is True and p['exchange']['ok'] and p['oracle_full_size']['ok'] and p['oracle_full_size']['max_rel'] < 1e-10
        assert p['oracle_subshard']['walkers'] >= 8 and p['veff_counts_per_bin']['ok']
        assert d['walker_sharded']['parity']['ok'] and d['e2e']['value'] != d['value'] and d['gpu_launches'] > 0
        if n > 1:
            assert d['weak']['sources_per_gpu'] == 10000000 and d['weak']['parity']['ok']
            assert d['config4']['sources_total'] == 100000000 and d['config4']['walkers'] == 2048 and d['config4']['parity']['ok']
    one = json.loads(open(os.path.join(ROOT, 'profiles', 'r02_bench_1gpu_final.json')).read().strip().splitlines()[-1])
    eight = json.loads(open(os.path.join(ROOT, 'profiles', 'r02_bench_8gpu.json')).read().strip().splitlines()[-1])
    assert eight['value'] > 1.0e12                               # the north_star target on its own configuration
    z = one['config2_z']                                         # BASELINE configs[2] rides in the default line
    assert z['sources_total'] == 1000000 and z['walkers'] == 512 and z['parity']['ok'] and 0.5 < z['frac_of_dfma_peak'] < 1.0
    assert 0.9 < eight['value'] / (8 * one['value']) <= 1.02     # strong-scaling efficiency
