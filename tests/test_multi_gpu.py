"""Real multi-GPU tests of the particle-sharded filter: one process per GPU (torchrun), the kernels
exchange over NVLink on their own.  Skipped on boxes with fewer than 2 GPUs; the ranks emulated on one
GPU (tests/test_gpu_parity.py, helpers.EmuRanks) cover the same code paths with a host-ordered exchange."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    from monte_carlo_localization_b200 import capi
    return capi.load_library().mcl_device_count()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torchrun(world, script, *args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, script)] + list(args)
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.parametrize("exchange,route", [("fused", "two-hop"), ("nccl", "two-hop"), ("fused", "one-hop")])
def test_sharded_filter_equals_single_gpu_bit_for_bit(exchange, route):
    """2 ranks x 131072 particles, 8 device-RNG updates + one update from degenerate weights (the overflow
    path of the exchange, every request going to one rank): indices, particles and weights bit-identical to
    the same filter on one GPU, for both routings of the resampling draws."""
    if _device_count() < 2:
        pytest.skip("needs 2 GPUs")
    res = _torchrun(2, "scripts/check_sharded_equals_single.py", "--particles-per-gpu", "131072", "--updates", "8",
                    "--exchange", exchange, "--route", route, "--degenerate")
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert res.returncode == 0 and lines, res.stdout[-2000:] + res.stderr[-4000:]
    r = json.loads(lines[-1])
    assert r["ok"] and r["indices_bit_identical"] and r["particles_bit_identical"] and r["weights_bit_identical"], r
    assert r["degenerate_particles_bit_identical"] and r["degenerate_weights_bit_identical"], r


def test_cpp_sharded_host_binary(tmp_path):
    """host/mcl_sharded: the C++ caller of mcl_create_sharded (one process per GPU, forked by the binary
    itself, the NCCL id handed over through pipes) tracks the ground truth and its ranks agree bit for bit."""
    import numpy as np
    from PIL import Image

    from monte_carlo_localization_b200 import maps
    if _device_count() < 2:
        pytest.skip("needs 2 GPUs")
    g = maps.load_named_map("sibal1")
    img = np.where(g.data[::-1] == 100, 0, np.where(g.data[::-1] == 0, 254, 205)).astype(np.uint8)
    Image.fromarray(img, "L").save(str(tmp_path / "sibal1.png"))
    y = str(tmp_path / "sibal1.yaml")
    with open(y, "w") as f:
        f.write("image: sibal1.png\nresolution: %r\norigin: [%r, %r, %r]\nnegate: 0\noccupied_thresh: 0.65\nfree_thresh: 0.1\n"
                % ((float(g.resolution),) + tuple(g.origin)))
    exe = os.path.join(ROOT, "monte_carlo_localization_b200", "host", "mcl_sharded")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(exe)])
    for extra in ([], ["--nccl-barrier"]):
        res = subprocess.run([exe, y, "--world", "2", "--particles", "65536", "--steps", "20", "--x", "-3.3", "--y", "1.6",
                              "--theta", "0.3"] + extra, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                             timeout=300, cwd=ROOT)
        lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
        assert res.returncode == 0 and lines, res.stdout[-3000:]
        r = json.loads(lines[-1])
        assert r["ranks_agree"] and r["failed_ranks"] == 0 and r["pose_error_m"] < 0.3, r
