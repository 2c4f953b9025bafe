"""
The reference arm of bench.py (`--impl reference`: the reference's own code -- /root/reference or the
staged copy baseline/_ref -- on all host cores, the numpy port as a second field) on the CPU:
rank 0 prints ONE JSON line with the contract's keys, every other rank exits 0 without work.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(rank):
    env = dict(os.environ, RANK=str(rank), WORLD_SIZE='2', LOCAL_RANK=str(rank), CUDA_VISIBLE_DEVICES='')
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                           '--steps', '1', '--warmup', '0', '--ref-positions', '16'],
                          env=env, capture_output=True, text=True, timeout=300)


def test_other_ranks_exit_without_work():
    r = run(1)
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_rank0_prints_the_contract_line():
    r = run(0)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'tested positions/sec (GMM+KS)' and d['unit'] == 'positions/s'
    assert d['higher_is_better'] is True and d['vs_baseline'] is None and d['n_gpus'] == 2
    assert d['value'] > 0 and d['steps'] == 1 and d['warmup'] == 0 and d['gpu_launches'] == 0
    cb = d['cpu_baseline']
    from oracle import refharness
    kind = 'reference' if refharness.reference_available() else 'port'
    assert cb['kind'] == kind and cb['cores'] == (os.cpu_count() or 1) and cb['value'] == d['value'] and cb['sample']
    if kind == 'reference':
        assert d['port']['kind'] == 'port' and d['port']['value'] > 0
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['config']['workload'].startswith('config3')


def test_product_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: without a CUDA device the product arm of bench.py stops with an error and
    prints no result line."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '0',
                        '--tx-per-batch', '1', '--no-cpu-baseline'], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
