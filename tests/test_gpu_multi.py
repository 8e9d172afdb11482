"""Multi-GPU parity (-m gpu; skipped on a single-GPU box): torch.distributed.run with one rank per GPU."""
import ctypes as C
import os
import socket
import subprocess
import sys

import pytest

import lowrankmodels_b200 as lrm

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    n = C.c_int32(0)
    lrm._abi.lib().glrmb200_device_count(C.byref(n))
    return n.value


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_fit_is_bit_identical_to_single_gpu(world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "mgpu_worker.py")], capture_output=True, text=True, timeout=900)
    open(os.path.join(ROOT, "gpurun_out", f"mgpu_worker_{world}.log"), "w").write(r.stdout + "\n----\n" + r.stderr) \
        if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else None
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("identical=True") == 6
