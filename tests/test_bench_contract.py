"""bench.py's contract lines that can be checked without a GPU: the reference arm (the reference's CPU algorithm through the
oracle port) prints ONE JSON line with the keys the driver reads, on the same metric / unit / workload string as the GPU arm,
and the GPU arm refuses to run without a device instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, cwd=ROOT, env=e, capture_output=True,
                          text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    r = _run(['--impl', 'reference', '--steps', '1', '--warmup', '0', '--no-configs'])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench
    assert d['impl'] == 'reference' and d['metric'] == bench.METRIC and d['unit'] == bench.UNIT
    assert d['config']['workload'] == bench.WORKLOAD['name']            # same_config with the GPU arm
    assert d['higher_is_better'] is True and d['dtype'] == 'f64' and d['data'] == 'synthetic' and d['vs_baseline'] is None
    assert d['value'] > 0 and d['steps'] == 1 and d['gpu_launches'] == 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': bench.UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(['--impl', 'reference', '--steps', '1', '--warmup', '0', '--no-configs'], env={'RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(['--steps', '1', '--warmup', '1', '--no-configs', '--no-cpu-baseline'])
    assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)
