"""GPU (>= 2 devices): island-partitioned population with global resampling over NVLink peer memory vs one GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import rcppsmc_b200 as smc
    return smc.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_island_matches_single_gpu(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs on one box (run under gpurun --gpus {world})")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "island_worker.py"), "18", "25"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    sys.stdout.write(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("-> OK") == 4
