"""Data-parallel parity on real GPUs: tools/dp_check.py under torchrun -- after one TrainStep on rank-sharded
data the all-reduced flat gradient and the parameter update equal those of one process over the concatenated batch.
Needs >= 2 GPUs (gpurun --gpus 2 / 8); the log of the last run is kept under profiles/."""
import os
import socket
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 8])
def test_dp_gradients_match_single_process(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                        os.path.join(ROOT, "tools", "dp_check.py")], capture_output=True, text=True, timeout=600, cwd=ROOT)
    out = r.stdout + r.stderr
    log_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(log_dir, exist_ok=True)
    with open(os.path.join(log_dir, f"r2_dp_check_{world}gpu.log"), "w") as f:
        f.write(out)
    assert r.returncode == 0 and "dp_check ok" in out, out[-3000:]
