"""bench.py contract on CPU: the reference arm (the only arm that runs without a GPU) prints ONE JSON line
with the keys the driver reads, for N = 1 and for the N > 1 launch (rank 0 prints, other ranks exit 0)."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1']
                       + extra, capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    return [l for l in r.stdout.splitlines() if l.startswith('{')]


def test_reference_arm_line_single_gpu():
    lines = _run([])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'cg_voxel_dof_per_s' and d['unit'] == 'voxel-DOF/s'
    assert d['higher_is_better'] is True and d['n_gpus'] == 1 and d['dtype'] == 'f64'
    assert d['value'] > 0 and abs(d['value']-d['e2e']['value']) < 1e-9*d['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    cb = d['cpu_baseline']
    # kind "reference" when the unmodified tree travelled under baseline/_ref, else the oracle port; the line names
    # the grid it actually timed (128^3) next to the GPU arm's grid
    assert cb['kind'] in ('port', 'reference') and cb['cores'] == 1 and cb['value'] == d['value'] and '128^3' in cb['sample']
    assert d['config']['grid'] == [128, 128, 128] and d['config']['gpu_arm_grid'] == [256, 256, 256]
    assert 'workload' in d['config'] and d['steps_timed'] == 1


def test_reference_arm_under_a_multi_rank_launch():
    lines = _run(['--gpus', '8'], env={'RANK': '0', 'WORLD_SIZE': '8', 'LOCAL_RANK': '0'})
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['n_gpus'] == 8 and d['config']['gpu_arm_grid'] == [512, 512, 512] and d['scaling'] == 'strong'
    assert _run(['--gpus', '8'], env={'RANK': '3', 'WORLD_SIZE': '8', 'LOCAL_RANK': '3'}) == []
