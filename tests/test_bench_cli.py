"""CPU: the reference arm of bench.py (`--impl reference`): the oracle port timed on the host cores, rank 0 only."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(extra_env, *args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='', **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', *args],
                          env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_other_ranks_exit_without_work():
    out = run_bench({'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'}, '--gpus', '2', '--steps', '1', '--warmup', '0')
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == ''


def test_reference_arm_prints_one_contract_line():
    out = run_bench({'RANK': '0', 'WORLD_SIZE': '1'}, '--steps', '1', '--warmup', '0')
    assert out.returncode == 0, out.stderr
    lines = [line for line in out.stdout.splitlines() if line.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line['impl'] == 'reference' and line['metric'].startswith('neurons described/sec')
    assert line['unit'] == 'neurons/s'
    for key in ('value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'gpu_launches'):
        assert key in line, key
    assert line['value'] > 0 and line['gpu_launches'] == 0
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e'] == {'value': line['value'], 'unit': line['unit'], 'h2d_bytes_per_step': 0,
                           'd2h_bytes_per_step': 0}
    assert line['config']['workload'] and 'model' not in line['config']
    sys.path.insert(0, ROOT)
    import bench
    assert line['config'] == bench.shared_config()  # both arms print the same dict (the driver's same_config check)
    assert line['arm']['device'] == 'cpu'
