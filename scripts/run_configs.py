#!/usr/bin/env python
"""Run the BASELINE.json configurations that are not the bench workload (parity-test cases and
functional checks) on the GPU box and print one JSON summary per config.

  config 1  sibal1, 4000 particles, single update (parity is in tests/test_gpu_parity.py)
  config 2  levine stand-in (basement_fixed), 100k particles, 1000-step replay, 1 GPU
  config 2s the same on the procedural levine stand-in (maps.synth_levine)
  config 4  1024 independent 4000-particle filters on sibal1 (one GPU's share of the batch
            is 128 filters; run at 1024 here to show the whole batch on one B200)
  config 5  global initialisation on the levine stand-in: particles uniform over free space,
            updates until the pose estimate stays within 0.25 m / 0.1 rad for 10 updates

Usage: python scripts/run_configs.py [--configs 1,2,4,5] [--n5 2000000]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from monte_carlo_localization_b200 import MclContext, maps, synth  # noqa: E402


def replay(ctx, grid, n_steps, speed, seed, batch=1, phase_stride=0):
    angles_full = synth.laser_angles()
    gt, actions = synth.trajectory(grid, n_steps + phase_stride * (batch - 1), speed)
    rng = np.random.default_rng(seed)
    obs = np.stack([synth.scan_from_pose(ctx.calc_range_many, gt[t + 1], angles_full, rng)[::18]
                    for t in range(len(actions))]).astype(np.float32)
    return gt, actions, obs


def timed_updates(ctx, actions, obs, gt, n_steps):
    errs = []
    ctx.synchronize()
    t0 = time.perf_counter()
    for t in range(n_steps):
        pose = ctx.update(actions[t], obs[t])
        errs.append(float(np.hypot(*(pose[:2] - gt[t + 1][:2]))))
    ctx.synchronize()
    return time.perf_counter() - t0, errs


def config1():
    g = maps.load_named_map("sibal1")
    ctx = MclContext(max_particles=4000, seed=20251)
    ctx.set_map(g)
    ctx.set_beam_angles(synth.beam_angles())
    gt, actions, obs = replay(ctx, g, 220, 3.0, 778)
    ctx.init_pose(gt[0])
    timed_updates(ctx, actions, obs, gt, 20)
    sec, errs = timed_updates(ctx, actions[20:], obs[20:], gt[20:], 200)
    return {"config": 1, "map": "sibal1", "particles": 4000, "beams": 60, "updates": 200, "ms_per_update": 1e3 * sec / 200,
            "updates_per_s": 200 / sec, "rays_per_s": 4000 * 60 * 200 / sec, "median_pose_err_m": float(np.median(errs))}


def config2(steps=1000, ray_mode=0, levine_synth=False):
    # levine.pgm is missing from the reference checkout: two stand-ins, the shipped
    # basement_fixed map and a procedural corridor loop honouring maps/levine.yaml
    g = maps.synth_levine() if levine_synth else maps.load_named_map("basement_fixed")
    N = 100000
    ctx = MclContext(max_particles=N, seed=20252)
    ctx.set_map(g)
    ctx.set_beam_angles(synth.beam_angles())
    ctx.set_ray_mode(ray_mode)
    gt, actions, obs = replay(ctx, g, steps + 20, 3.0, 779)
    ctx.init_pose(gt[0])
    timed_updates(ctx, actions, obs, gt, 20)
    sec, errs = timed_updates(ctx, actions[20:], obs[20:], gt[20:], steps)
    return {"config": 2, "map": ("levine_synth (procedural stand-in honouring maps/levine.yaml)" if levine_synth else
                                 "basement_fixed (levine stand-in: levine.pgm is missing from the reference checkout)"),
            "particles": N, "beams": 60, "updates": steps, "ms_per_update": 1e3 * sec / steps, "updates_per_s": steps / sec,
            "rays_per_s": N * 60 * steps / sec, "median_pose_err_m": float(np.median(errs)),
            "max_pose_err_m": float(np.max(errs)), "ray_mode": ray_mode, "ray_stage": ctx.ray_stage_info()}


def config4(F=1024, steps=20, ray_mode=0):
    g = maps.load_named_map("sibal1")
    N = 4000
    ctx = MclContext(max_particles=N, num_filters=F, seed=20254)
    ctx.set_map(g)
    ctx.set_beam_angles(synth.beam_angles())
    ctx.set_ray_mode(ray_mode)
    stride = 1
    gt, actions, obs = replay(ctx, g, steps + 5, 3.0, 781, batch=F, phase_stride=stride)
    for f in range(F):   # every car starts at its own phase of the lap
        ctx.init_pose(gt[f * stride], filter=f)

    def step(t):
        idx = np.arange(F) * stride + t
        return ctx.update(actions[idx], obs[idx]), idx

    for t in range(5):
        step(t)
    ctx.synchronize()
    t0 = time.perf_counter()
    for t in range(5, 5 + steps):
        poses, idx = step(t)
    ctx.synchronize()
    sec = time.perf_counter() - t0
    err = np.hypot(poses[:, 0] - gt[idx + 1, 0], poses[:, 1] - gt[idx + 1, 1])
    return {"config": 4, "map": "sibal1", "filters": F, "particles_per_filter": N, "beams": 60, "batch_steps": steps,
            "ms_per_batch_step": 1e3 * sec / steps, "filter_updates_per_s": F * steps / sec,
            "rays_per_s": F * N * 60 * steps / sec, "median_pose_err_m": float(np.median(err)),
            "frac_filters_within_0.3m": float(np.mean(err < 0.3)), "ray_mode": ray_mode, "ray_stage": ctx.ray_stage_info()}


def config5(N=2000000, max_updates=60, ray_mode=0):
    g = maps.load_named_map("basement_fixed")
    ctx = MclContext(max_particles=N, seed=20255)
    ctx.set_map(g)
    ctx.set_beam_angles(synth.beam_angles())
    ctx.set_ray_mode(ray_mode)
    gt, actions, obs = replay(ctx, g, max_updates, 3.0, 782)
    ctx.init_global()
    streak, conv, times = 0, None, []
    for t in range(max_updates):
        t0 = time.perf_counter()
        pose = ctx.update(actions[t], obs[t])
        times.append(time.perf_counter() - t0)
        d = np.hypot(*(pose[:2] - gt[t + 1][:2]))
        dth = abs((pose[2] - gt[t + 1][2] + np.pi) % (2 * np.pi) - np.pi)
        streak = streak + 1 if (d < 0.25 and dth < 0.1) else 0
        if streak == 10 and conv is None:
            conv = t + 1 - 9
    return {"config": 5, "map": "basement_fixed (levine stand-in)", "particles": N, "beams": 60,
            "free_cells": ctx.num_free_cells(), "updates_run": max_updates,
            "converged_at_update": conv, "ms_per_update_first5": 1e3 * float(np.mean(times[:5])),
            "ms_per_update_last5": 1e3 * float(np.mean(times[-5:])),
            "time_to_converge_s": None if conv is None else float(np.sum(times[:conv + 9])),
            "final_pose_err_m": float(d), "ray_mode": ray_mode, "ray_stage": ctx.ray_stage_info()}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,4,5")
    ap.add_argument("--n5", type=int, default=2000000)
    ap.add_argument("--steps2", type=int, default=1000)
    ap.add_argument("--filters4", type=int, default=1024)
    ap.add_argument("--ray-mode", type=int, default=0, help="0 auto, 1 isotropic kernel only, 2 directional stage always")
    a = ap.parse_args()
    for c in a.configs.split(","):
        if c == "2s":
            print(json.dumps(config2(a.steps2, a.ray_mode, levine_synth=True)))
        elif c == "1":
            print(json.dumps(config1()))
        elif c == "2":
            print(json.dumps(config2(a.steps2, a.ray_mode)))
        elif c == "4":
            print(json.dumps(config4(a.filters4, ray_mode=a.ray_mode)))
        elif c == "5":
            print(json.dumps(config5(a.n5, ray_mode=a.ray_mode)))
        sys.stdout.flush()
