"""world_size-2 gloo test (CPU) of the particle-sharding HOST logic in
monte_carlo_localization_b200/sharded.py: the slot plan, the bootstrap that carries the NCCL id from
rank 0 to the others, and the slicing / gathering of whole-filter arrays.  The per-slice compute is
done with the oracle here; the protocol (own slice of the global multinomial draw -> motion -> weights
-> global sequential sum) must reproduce the single-filter oracle update bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import load_golden
from monte_carlo_localization_b200 import maps
from monte_carlo_localization_b200.sharded import ShardPlan, bootstrap_id, gather_host, local_slice
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
        # the id is made on rank 0 only and must arrive unchanged everywhere
        made = bytes(range(128))
        got = bootstrap_id(lambda: made, rank)
        assert got == made
        z = load_golden("update_sibal1_4000.npz")
        g = maps.load_named_map("sibal1")
        N = int(z["N"])
        plan = ShardPlan(N, world)
        lo, cnt = plan.slots(rank)
        P, W = z["init_particles"].copy(), z["init_weights"].copy()   # whole filter, for the oracle's global CDF
        slice_orc = ob.Oracle(g, z["angles"], max_particles=cnt)      # per-slice compute stand-in
        for t in range(len(z["u"])):
            u, zz = z["u"][t], z["z"][t]
            # own slots of the GLOBAL multinomial draw, motion and weights of the own slice
            idx = ob.resample_indices(W, u[lo:lo + cnt])
            _, z_loc = local_slice(plan, rank, weights=zz, per_particle=3)
            prop = slice_orc.motion_model(P[:, idx], z["actions"][t], z_loc)
            w_loc = slice_orc.sensor_weights(prop, z["obs"][t])
            # whole filter from the slices
            P, w_raw = gather_host(plan, prop, w_loc)
            s = np.add.accumulate(w_raw)[-1]       # the reference's sequential sum (:679)
            W = w_raw / s
            assert np.array_equal(P, z["particles"][t]), "rank %d t %d" % (rank, t)
            assert np.array_equal(W, z["weights"][t]), "rank %d t %d weights" % (rank, t)
            p_loc, w_sl = local_slice(plan, rank, P, W)
            assert np.array_equal(p_loc, prop) and np.array_equal(w_sl, W[lo:lo + cnt])
        np.save(os.path.join(out_dir, "w_rank%d.npy" % rank), W)
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
