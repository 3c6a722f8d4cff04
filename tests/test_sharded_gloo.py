"""world_size-2 gloo test (CPU) of the particle-sharding host logic: the slot plan and the
in-place exchange that monte_carlo_localization_b200/sharded.py uses on NCCL.  The per-slice
compute is done with the oracle here; the protocol (own slots -> exchange -> global finish)
must reproduce the single-filter oracle update bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import load_golden
from monte_carlo_localization_b200 import maps
from monte_carlo_localization_b200.sharded import ShardPlan, exchange
from oracle import bindings as ob


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z = load_golden("update_sibal1_4000.npz")
        g = maps.load_named_map("sibal1")
        N = int(z["N"])
        plan = ShardPlan(N, world)
        lo, cnt = plan.slots(rank)
        # full state on every rank
        x, y, th = (torch.from_numpy(z["init_particles"][k].copy()) for k in range(3))
        w = torch.from_numpy(z["init_weights"].copy())
        slice_orc = ob.Oracle(g, z["angles"], max_particles=cnt)   # per-slice compute stand-in
        for t in range(len(z["u"])):
            u, zz = z["u"][t], z["z"][t]
            # --- local: resample own slots from the GLOBAL cdf, motion, weights -------------
            idx = ob.resample_indices(w.numpy(), u[lo:lo + cnt])
            prop = np.stack([x.numpy()[idx], y.numpy()[idx], th.numpy()[idx]])
            prop = slice_orc.motion_model(prop, z["actions"][t], zz[3 * lo:3 * (lo + cnt)])
            w_loc = slice_orc.sensor_weights(prop, z["obs"][t])
            nx, ny, nth, nw = x.clone(), y.clone(), th.clone(), w.clone()
            nx[lo:lo + cnt] = torch.from_numpy(prop[0])
            ny[lo:lo + cnt] = torch.from_numpy(prop[1])
            nth[lo:lo + cnt] = torch.from_numpy(prop[2])
            nw[lo:lo + cnt] = torch.from_numpy(w_loc)
            # --- exchange ---------------------------------------------------------------------
            exchange([nx, ny, nth, nw], plan, rank)
            # --- finish: sequential global sum, normalise --------------------------------------
            s = np.add.accumulate(nw.numpy())[-1]
            w = nw / s
            x, y, th = nx, ny, nth
            assert np.array_equal(np.stack([x.numpy(), y.numpy(), th.numpy()]), z["particles"][t]), "rank %d t %d" % (rank, t)
            assert np.array_equal(w.numpy(), z["weights"][t]), "rank %d t %d weights" % (rank, t)
        np.save(os.path.join(out_dir, "w_rank%d.npy" % rank), w.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_protocol_equals_single_filter(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    w0 = np.load(tmp_path / "w_rank0.npy")
    w1 = np.load(tmp_path / "w_rank1.npy")
    assert np.array_equal(w0, w1)   # ranks stay bit-identical


def test_shard_plan():
    p = ShardPlan(4000, 8)
    assert p.n_local == 500 and p.slots(3) == (1500, 500)
    assert p.owner([0, 499, 500, 3999]).tolist() == [0, 0, 1, 7]
    with pytest.raises(ValueError):
        ShardPlan(4001, 8)
    with pytest.raises(ValueError):
        p.slots(8)
