"""Particle-sharded operation of ONE global filter: one process per GPU.

The reference has a single process and no collectives.  Its update couples particles in three
places -- the weight sum (src/particle_filter.cpp:679), the global CDF + source gather of the
multinomial resampling (:658-665) and the expected-pose sums (:702-710).  Sharding keeps the
reference's semantics exactly:

  * every rank holds the whole filter state, but computes only output slots
    [rank * n_local, (rank + 1) * n_local) of the expensive per-particle stages
    (resample search, motion, ray cast, weights);
  * ONE exchange step per update over NCCL / NVLink: in mode "p2p" an in-place all-gather of
    the raw weights (+ four pose partial sums per rank), the source poses being read by the
    resampling kernel straight from their owner's memory (CUDA IPC); in mode "allgather" an
    in-place all-gather of all four state arrays (x, y, theta, raw weight);
  * the global weight sum, normalisation, pose and the next CDF are then computed on every
    rank from identical data with deterministic kernels, so all ranks stay bit-identical and
    the gathered result equals the single-filter update with the same noise.

`ShardPlan` and `exchange` are backend-agnostic host logic (exercised with gloo on CPU in
tests/test_sharded_gloo.py); `ShardedFilter` binds them to the CUDA context.
scripts/check_sharded_equals_single.py checks on real GPUs that the sharded filter stays
bit-identical to the same filter on one GPU.
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


def exchange(arrays, plan: ShardPlan, rank: int, group=None):
    """All-gather, in place, the rank's slice of every array in `arrays` (1-D torch tensors of
    length n_global living on the backend's device)."""
    import torch.distributed as dist
    lo, cnt = plan.slots(rank)
    backend = dist.get_backend(group)
    for full in arrays:
        if full.numel() != plan.n_global:
            raise ValueError("array of %d elements, expected %d" % (full.numel(), plan.n_global))
        if backend == "nccl":
            # in place: the send buffer is the rank's own slice of the receive buffer
            dist.all_gather_into_tensor(full, full[lo:lo + cnt], group=group)
        else:
            mine = full[lo:lo + cnt].clone()
            dist.all_gather([full[r * cnt:(r + 1) * cnt] for r in range(plan.world)], mine, group=group)


class _DevArray:
    """A device pointer dressed as a __cuda_array_interface__ object so torch can alias it."""

    def __init__(self, ptr: int, n: int, typestr: str = "<f8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class ShardedFilter:
    """One rank's share of a particle-sharded global filter (needs torch.distributed + NCCL)."""

    def __init__(self, grid, angles, n_local: int, rank: int, world: int, device: int = 0, seed: int = 0,
                 mode: str = "p2p", **params):
        """mode "allgather": all four state arrays are all-gathered every update (32 B/particle).
        mode "p2p": ranks map each other's state arrays with CUDA IPC; the resampling kernel reads
        source poses from their owner over NVLink and only raw weights + 4 pose partial sums per
        rank are all-gathered (8 B/particle)."""
        import torch
        from .capi import MclContext
        if mode not in ("allgather", "p2p"):
            raise ValueError("mode must be 'allgather' or 'p2p'")
        self.torch = torch
        self.rank, self.world, self.device, self.mode = rank, world, device, mode
        self.plan = ShardPlan(n_local * world, world)

        def make_ctx():
            c = MclContext(device=device, max_particles=self.plan.n_global, seed=seed, **params)
            c.set_map(grid)
            c.set_beam_angles(angles)
            c.set_shard(*self.plan.slots(rank))
            return c

        self.ctx = make_ctx()
        self._alias = {}
        self.p2p_error = None
        if mode == "p2p":
            import torch.distributed as dist
            ok = 1
            try:
                blobs = [None] * world
                dist.all_gather_object(blobs, self.ctx.ipc_export())
                self.ctx.ipc_import(world, rank, b"".join(blobs))
            except Exception as e:   # e.g. peer access or CUDA IPC not permitted on this host
                ok, self.p2p_error = 0, str(e)
            flag = torch.tensor([ok], dtype=torch.int32, device="cuda:%d" % device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                # every rank falls back together: all four state arrays are all-gathered instead
                self.mode = "allgather"
                self.ctx.close()
                self.ctx = make_ctx()
            dist.barrier()

    def init_pose(self, pose, normals_3n=None):
        # every rank initialises the full state; the device RNG is keyed by the global slot,
        # so all ranks hold identical particles
        self.ctx.init_pose(pose, normals_3n)

    def init_global(self):
        """initialize_global (src/particle_filter.cpp:401-446) with the device RNG: keyed by the
        global slot, so every rank holds the same particles."""
        self.ctx.init_global()

    def _tensor(self, ptr: int, n: int):
        t = self._alias.get(ptr)
        if t is None:
            t = self.torch.as_tensor(_DevArray(ptr, n), device="cuda:%d" % self.device)
            self._alias[ptr] = t
        return t

    def update_dev(self, action_dev_ptr: int, obs_dev_ptr: int, u_dev_ptr: int = 0, z_dev_ptr: int = 0):
        """One MCL update; inputs on the device; the ctx must launch on torch's current stream."""
        import torch.distributed as dist
        self.ctx.update_local_dev(action_dev_ptr, obs_dev_ptr, u_dev_ptr, z_dev_ptr)
        if self.mode == "p2p":
            w_ptr, part_ptr = self.ctx.p2p_buffers_dev()
            exchange([self._tensor(w_ptr, self.plan.n_global)], self.plan, self.rank)
            part = self._tensor(part_ptr, 4 * self.world)
            dist.all_gather_into_tensor(part, part[4 * self.rank:4 * self.rank + 4])
        else:
            ptrs, n, _, _ = self.ctx.exchange_buffers_dev()
            exchange([self._tensor(p, n) for p in ptrs], self.plan, self.rank)
        self.ctx.update_finish_dev()

    def gather_state(self):
        """(particles [3, N], weights [N]) of the whole filter on every rank (collective)."""
        import torch.distributed as dist
        torch = self.torch
        p = torch.from_numpy(self.ctx.get_particles()).cuda(self.device)
        if self.mode == "p2p":   # only the own slice of the poses is current
            lo, cnt = self.plan.slots(self.rank)
            for k in range(3):
                row = p[k].contiguous()
                dist.all_gather_into_tensor(row, row[lo:lo + cnt].clone())
                p[k] = row
        return p.cpu().numpy(), self.ctx.get_weights()

    def set_state(self, particles, weights):
        self.ctx.set_particles(particles, weights)

    def update(self, action, obs, u=None, z3n=None):
        """Host-facing update: action/scan (and optional injected noise for the WHOLE filter)
        copied in, pose copied out."""
        torch = self.torch
        a = torch.as_tensor(np.ascontiguousarray(action, dtype=np.float64)).cuda(self.device)
        o = torch.as_tensor(np.ascontiguousarray(obs, dtype=np.float32)).cuda(self.device)
        ud = None if u is None else torch.as_tensor(np.ascontiguousarray(u, dtype=np.float64)).cuda(self.device)
        zd = None if z3n is None else torch.as_tensor(np.ascontiguousarray(z3n, dtype=np.float64)).cuda(self.device)
        self.update_dev(a.data_ptr(), o.data_ptr(), 0 if ud is None else ud.data_ptr(), 0 if zd is None else zd.data_ptr())
        return self.ctx.read_pose()
