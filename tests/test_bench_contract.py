"""bench.py contract checks that need no GPU: the reference arm (the CPU oracle timed on the host cores) prints ONE
JSON line with the keys the driver reads, ranks other than 0 stay silent, and the product path refuses to run
without CUDA instead of falling back to the oracle."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, cwd=ROOT, env=e, capture_output=True,
                          text=True, timeout=600)


def test_reference_arm_line():
    r = _run(['--impl', 'reference', '--steps', '1', '--warmup', '0'])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['unit'] == 'ref-views/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['steps'] == 1 and d['vs_baseline'] is None and d['data'] == 'synthetic'
    assert d['config']['workload'].startswith('C2') and 'model' not in d['config']
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value'] == d['e2e']['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0


def test_reference_arm_other_ranks_exit_silently():
    r = _run(['--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'],
             env={'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [l for l in r.stdout.splitlines() if l.startswith('{')]


def test_product_arm_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    r = _run(['--steps', '1', '--warmup', '0'])
    assert r.returncode != 0                                  # no CPU fallback, no oracle behind the product path
    assert not [l for l in r.stdout.splitlines() if l.startswith('{')]


def test_graft_entry_builds_and_smoke_needs_a_gpu():
    """build() is the driver's CPU-side "does it build" check; smoke() must refuse to run without cuda:0 rather
    than check the oracle against itself"""
    import importlib
    import torch
    sys.path.insert(0, ROOT)
    g = importlib.import_module('__graft_entry__')
    g.build()
    assert os.path.exists(os.path.join(ROOT, '3dvnet_b200', 'lib3dvnet_b200.so'))
    if not torch.cuda.is_available():
        with pytest.raises(AssertionError, match='cuda:0'):
            g.smoke()
