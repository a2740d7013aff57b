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
