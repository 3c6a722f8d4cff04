"""Particle-sharded operation of ONE global filter: one process per GPU.

The reference has a single process and no collectives.  Its update couples particles in three
places -- the weight sum (src/particle_filter.cpp:679), the CDF + source gather of the multinomial
resampling (:658-665) and the expected-pose sums (:702-710).  The sharding lives in the C library
(`mcl_create_sharded`, include/mcl_b200.h): every rank holds only its own slot range of every
per-particle array, the three sequentially rounded reductions exchange a < 2 KB summary per rank,
resampling is sender-driven (the slot's owner sends a 4-byte request to the rank whose CDF range holds
the draw, that rank pushes the source pose back over NVLink), and the kernels perform the exchanges themselves -- an update has no host call
between its launches.  With the same noise the ranks' slices equal the single-GPU filter bit for
bit (scripts/check_sharded_equals_single.py, tests/test_multi_gpu.py).

This module is the thin Python caller: `torch.distributed` only carries the 128-byte NCCL id from
rank 0 to the others (any bootstrap would do); the library owns the communicator.  `ShardPlan`,
`bootstrap_id`, `local_slice` and `gather_host` are backend-agnostic host logic, exercised with gloo
on CPU in tests/test_sharded_gloo.py.
"""
from __future__ import annotations

import numpy as np


class ShardPlan:
    """Contiguous, equal slot ranges: rank r owns [r * n_local, (r + 1) * n_local)."""

    def __init__(self, n_global: int, world: int):
        if world < 1 or n_global < world:
            raise ValueError("bad shard plan: %d particles over %d ranks" % (n_global, world))
        if n_global % world:
            raise ValueError("max_particles (%d) must be a multiple of the number of ranks (%d)" % (n_global, world))
        self.n_global = n_global
        self.world = world
        self.n_local = n_global // world

    def slots(self, rank: int):
        if not 0 <= rank < self.world:
            raise ValueError("rank %d outside [0,%d)" % (rank, self.world))
        return rank * self.n_local, self.n_local

    def owner(self, slot):
        return np.asarray(slot) // self.n_local


def bootstrap_id(make_id, rank: int, group=None) -> bytes:
    """Rank 0 calls make_id() (mcl_nccl_unique_id) and every rank receives the bytes."""
    import torch.distributed as dist
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    return bytes(box[0])


def local_slice(plan: ShardPlan, rank: int, particles=None, weights=None, per_particle: int = 1):
    """The rank's share of whole-filter arrays: particles [3, NG] -> [3, n]; weights [NG * k] -> [n * k]."""
    lo, cnt = plan.slots(rank)
    p = None if particles is None else np.ascontiguousarray(np.asarray(particles).reshape(3, plan.n_global)[:, lo:lo + cnt])
    w = None if weights is None else np.ascontiguousarray(
        np.asarray(weights).reshape(-1)[lo * per_particle:(lo + cnt) * per_particle])
    return p, w


def gather_host(plan: ShardPlan, local_particles, local_weights, group=None):
    """Whole-filter (particles [3, NG], weights [NG]) from every rank's host slices, over any backend."""
    import torch
    import torch.distributed as dist
    mine = torch.from_numpy(np.concatenate([np.asarray(local_particles).reshape(3, plan.n_local),
                                            np.asarray(local_weights).reshape(1, plan.n_local)]).copy())
    parts = [torch.empty_like(mine) for _ in range(plan.world)]
    dist.all_gather(parts, mine, group=group)
    full = torch.cat(parts, dim=1).numpy()
    return np.ascontiguousarray(full[:3]), np.ascontiguousarray(full[3])


class ShardedFilter:
    """One rank's share of a particle-sharded global filter (torch.distributed carries the NCCL id)."""

    def __init__(self, grid, angles, n_local: int, rank: int, world: int, device: int = 0, seed: int = 0,
                 exchange: str = "fused", route: str = "auto", **params):
        """exchange "fused": the kernels publish and wait on their own (NVLink stores + system-scope
        flags).  "nccl": the same stores, but the ranks meet in a one-word ncclAllGather enqueued
        between the publishing and the consuming kernel (for comparison).
        route "two-hop": request routing, work per rank independent of the world size; "one-hop": every
        rank tests all draws; "auto" (default): the library's choice, two hops from 3 ranks on
        (mcl_shard_set_route)."""
        from . import capi
        if exchange not in ("fused", "nccl"):
            raise ValueError("exchange must be 'fused' or 'nccl'")
        self.rank, self.world, self.device, self.exchange = rank, world, device, exchange
        self.plan = ShardPlan(n_local * world, world)
        nccl_id = bootstrap_id(capi.nccl_unique_id, rank)
        self.ctx = capi.MclContext(device=device, shard=(world, rank), nccl_id=nccl_id,
                                   max_particles=self.plan.n_global, seed=seed, **params)
        self.ctx.set_map(grid)
        self.ctx.set_beam_angles(angles)
        if exchange == "nccl":
            self.ctx.shard_set_exchange(False)
        if route not in ("auto", "two-hop", "one-hop"):
            raise ValueError("route must be 'auto', 'two-hop' or 'one-hop'")
        self.route = route if route != "auto" else ("two-hop" if world >= 3 else "one-hop")
        if route != "auto":
            self.ctx.shard_set_route(route == "two-hop")

    def init_pose(self, pose, normals_3n=None):
        """initialize_particles_pose of the WHOLE filter; the device RNG is keyed by the global slot.
        normals_3n (optional): the injected normals of the whole filter."""
        z = None
        if normals_3n is not None:
            _, z = local_slice(self.plan, self.rank, weights=normals_3n, per_particle=3)
        self.ctx.init_pose(pose, z)

    def init_global(self):
        """initialize_global (src/particle_filter.cpp:401-446), device RNG keyed by the global slot."""
        self.ctx.init_global()

    def update_dev(self, action_dev_ptr: int, obs_dev_ptr: int):
        self.ctx.update_dev(action_dev_ptr, obs_dev_ptr)

    def update(self, action, obs, u=None, z3n=None):
        """Host-facing update: action/scan (and optional injected noise for the WHOLE filter)
        copied in, pose of the whole filter copied out."""
        return self.ctx.update(action, obs, u, z3n)

    def gather_state(self):
        """(particles [3, NG], weights [NG]) of the whole filter on every rank (NCCL, collective)."""
        return self.ctx.sharded_gather()

    def set_state(self, particles, weights):
        p, w = local_slice(self.plan, self.rank, particles, weights)
        self.ctx.set_particles(p, w)

    def close(self):
        self.ctx.close()
