"""Row-partitioned multi-GPU solve (needs >= 2 GPUs on the box; skipped otherwise)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize('world', [2, 4])
def test_distributed_cg_matches_single_gpu(world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
           '--master-addr', '127.0.0.1', '--master-port', str(29600 + world),
           os.path.join(ROOT, 'tools', 'dd_bench.py'), '--refine', '0', '--h', '0.04', '--reps', '2',
           '--replicate-below', '3000']
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith('{')][-1]
    out = json.loads(line)
    assert out['distributed']['converged'] and out['single_gpu']['converged']
    assert out['rel_l2_vs_single_gpu'] < 1e-10          # parity bar of the path (fields 1e-10 relative L2)
    assert abs(out['distributed']['iterations'] - out['single_gpu']['iterations']) <= 2


@pytest.mark.parametrize('world', [2])
def test_distributed_full_step_matches_single_gpu(world):
    """BASELINE config 5 through bench.py's dd_strong leg: Taylor-Hood MINRES + adv-diff FGMRES row-partitioned over
    the ranks (peer-memory halos, in-kernel all-reduces) must reproduce the single-GPU fields to 1e-10."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
           '--master-addr', '127.0.0.1', '--master-port', str(29700 + world),
           os.path.join(ROOT, 'bench.py'), '--gpus', str(world), '--h', '0.04', '--refine', '0', '--steps', '2',
           '--warmup', '1', '--no-cpu', '--dd-replicate-below', '3000', '--dd-refine', '']
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith('{')][-1]
    out = json.loads(line)
    dd = out['dd_strong'][0]
    assert len(out['dd_strong']) == 1 and dd['n_gpus'] == world and dd['parity_ok'], dd
    assert dd['local_assembly_equals_replicated']             # distributed assembly == rows of the replicated one, bit for bit
    assert dd['distributed_levels']['velocity'] >= 2          # a multi-level row-partitioned hierarchy was exercised
    it1, itn = dd['iterations']['single'], dd['iterations']['distributed']
    assert abs(it1['stokes'] - itn['stokes']) <= 3 and abs(it1['advdiff'] - itn['advdiff']) <= 2
