"""Multi-GPU row slabs: bit-identical to one GPU.  Needs >= 2 GPUs on the box (skipped on a
one-GPU box); the host-side arithmetic is covered on CPU by tests/test_slab_cpu.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.gpu
@pytest.mark.parametrize("world,opts", [(2, ""), (4, ""), (2, "sor_slab_inpass=1")])
def test_slab_run_equals_one_gpu(world, opts):
    """opts: library options of the worker (W2_OPTS) -- the in-pass edge stores of the fused SOR are the non-default
    variant and get their own run."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29610 + world + (10 if opts else 0)),
           os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    env = dict(os.environ, W2_OPTS=opts)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    sys.stdout.write(r.stdout[-6000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0, "slab run differs from the one-GPU run (see output)"
