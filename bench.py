#!/usr/bin/env python
"""Headline benchmark: rodent-imitation env (control) steps / second at 8192 envs per B200 (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--envs E] [--model rodent]

Own arm: one "step" = ONE launch of the fused kernel (5 physics substeps + tracking reward + observation +
episode / auto-reset bookkeeping) over all envs of the rank.  `value` = env-steps/s with the state resident in HBM
(actions pre-generated on the device), CUDA-event timed per launch with an L2 flush between launches; `e2e` = the
same through the public env API with HOST buffers (pinned action H2D + obs/reward/done D2H inside the timed region).
`roofline` = algorithmic HBM bytes (4776 B / env-step for the rodent) / launch time against the
measured copy bandwidth.  `cpu_baseline` / `--impl reference` = the C restatement of the reference path ("port":
the reference itself is Python on un-vendored jax / mujoco-mjx / brax, none installable here) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "rodent-imitation env-steps/sec at 8192 envs per GPU"
UNIT = "env-steps/s"
# algorithmic HBM bytes per env-step (SURVEY.md section 8d / BASELINE.md section 2):
#   read  qpos + qvel + act + qacc_warmstart + time + action + 2 ints
#   write qpos + qvel + act + qacc_warmstart + time + obs + reward, done + 12 metrics + 3 info floats + 2 info ints
ALGO_BYTES = dict(rodent=4776, fly_free=6472, fly_tethered=6128, rodent_pair=9268)
MODEL_XML = dict(rodent="rodent", fly_free="fruitfly_force_fast (free root)", fly_tethered="fruitfly_force_fast (tethered)",
                 rodent_pair="rodent_pair")


def algo_bytes(nq, nv, na, nu, obs):
    return 4 * ((nq + 2 * nv + na + 1 + nu + 2) + (nq + 2 * nv + na + 1 + obs + 2 + 12 + 3 + 2))


def workload_name(model, envs):
    return f"{MODEL_XML[model]}.xml imitation rollout, {envs} envs/GPU, n_frames=5 substeps, CG 4x4, step-only"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(model):
    """Per-launch dram bytes of the step kernel from the committed ncu capture, if one exists."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(model)
    return None


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_port_run(model, n_envs, n_steps, threads):
    """Times the oracle port (C restatement of mjx.step x n_frames + numpy env layer) on `threads` host threads."""
    import common
    import oracle as oracle_mod
    import env_oracle
    m, cfg, clip, _ = common.setup(model)
    o = oracle_mod.Oracle(m, np.float32)
    eo = env_oracle.EnvOracle(o, clip, cfg, dtype=np.float32)
    keys = common.jax_keys(n_envs)
    s = eo.reset(keys)
    acts = common.actions(n_steps + 1, n_envs, m.nu, seed=1)
    # warm-up step (thread pool + caches)
    s = eo.step(s, acts[0])
    t0 = time.perf_counter()
    for t in range(n_steps):
        s = eo.step(s, acts[t + 1])
    dt = time.perf_counter() - t0
    return n_envs * n_steps / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_envs = 64 * cores if args.ref_envs is None else args.ref_envs
    # the real MJX env on the JAX CPU backend when it is importable (baseline/_ref) ...
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import run_cpu_baseline as rcb
    real = rcb.run_mjx(min(n_envs, 1024), args.steps) if args.model == "rodent" else None
    if real is not None:
        v, dt, kind, sample, cores = real["value"], real["wall_s"], "reference", real["sample"], real["cores"]
    else:
        # ... else the oracle port.  Every "step" is a bounded sample: one control step over n_envs envs on all host threads
        for _ in range(args.warmup):
            cpu_port_run(args.model, min(n_envs, 2 * cores), 1, cores)
        v, dt = cpu_port_run(args.model, n_envs, args.steps, cores)
        kind = "port"
        sample = f"{n_envs} envs x {args.steps} control steps, oracle port (C, float32), {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.model, args.envs), "sample": sample,
                   "same_config": False, "note": "a bounded sample of the per-environment-linear CPU path, not all 8192 environments"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from brax_tracking_b200 import envs, native, parallel, presets, prng

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    saved_stdout = None
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its version banner to fd 1 when the communicator is created, so
        # fd 1 points at stderr until the line is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    n, K, Wm = args.envs, args.steps, args.warmup
    base = presets.make_env(args.model, device=local)
    m = base.sys
    env = envs.wrap(base, episode_length=base.episode_length)
    assert algo_bytes(m.nq, m.nv, m.na, m.nu, env.observation_size) == ALGO_BYTES[args.model]
    # per-rank env shard: keys = split(PRNGKey(0), world*n)[rank*n:(rank+1)*n]  (custom_ppo.py:220-223)
    keys = parallel.shard_keys(prng.PRNGKey(0), world * n, rank, world)
    state = env.reset(keys)
    rng = np.random.default_rng(1 + rank)
    n_act = 8
    acts = torch.from_numpy(np.tanh(rng.standard_normal((n_act, n, m.nu))).astype(np.float32)).to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(Wm):
        state = env.step(state, acts[i % n_act])
    barrier()
    # ---------------- device-resident timing: one CUDA-event pair per launch, L2 flushed between launches
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = native.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    for i in range(K):
        flush.zero_()
        ev[i][0].record()
        state = env.step(state, acts[i % n_act])
        ev[i][1].record()
    barrier()
    launches = native.launch_count() - launches0
    kernel_ms = [a.elapsed_time(b) for a, b in ev]
    t_dev = sum(kernel_ms) / 1e3
    # ---------------- end to end through the public API with HOST buffers
    h_act = torch.from_numpy(np.tanh(rng.standard_normal((n_act, n, m.nu))).astype(np.float32)).pin_memory()
    d_act = torch.empty(n, m.nu, dtype=torch.float32, device=dev)
    h_rew = torch.empty(n, dtype=torch.float32).pin_memory()
    h_done = torch.empty(n, dtype=torch.float32).pin_memory()
    # observations of a host-side consumer: the kernel writes each env's row straight into page-locked host memory
    # (mapped pointer; the device->host transfer of the 20 MB of observations overlaps the launch)
    h_obs = env.bind_host_obs(state)
    for i in range(2):
        d_act.copy_(h_act[i % n_act], non_blocking=True)
        state = env.step(state, d_act)
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        d_act.copy_(h_act[i % n_act], non_blocking=True)
        state = env.step(state, d_act)          # obs rows land in h_obs (host) during the launch
        h_rew.copy_(state.reward, non_blocking=True)
        h_done.copy_(state.done, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the host consumes obs/reward/done every step
    e1.record()
    barrier()
    t_e2e = max(e0.elapsed_time(e1) / 1e3, 0.0)
    t_wall = time.perf_counter() - t0
    sampler.stop_flag = True
    sampler.join()
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(tt[0]), float(tt[1])
    # ---------------- the two measurements that involve the learner / a fixed global batch, same JSON line:
    #   strong: the SAME 8192 environments split over the N ranks (step only, no collective: per-rank launches shrink)
    #   train : custom_ppo.train on 8192 envs per GPU with the train_fly.yaml hyper-parameters; NCCL gradient all-reduce per minibatch
    strong = train_blk = None
    if not args.no_extra:
        del state, h_obs
        gn = 8192
        if gn % world == 0:
            ns = gn // world
            st2 = env.reset(parallel.shard_keys(prng.PRNGKey(0), gn, rank, world))
            a2 = torch.from_numpy(np.tanh(rng.standard_normal((n_act, ns, m.nu))).astype(np.float32)).to(dev)
            for i in range(Wm):
                st2 = env.step(st2, a2[i % n_act])
            ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
            barrier()
            for i in range(K):
                flush.zero_()
                ev2[i][0].record()
                st2 = env.step(st2, a2[i % n_act])
                ev2[i][1].record()
            barrier()
            ts_ = torch.tensor([sum(a.elapsed_time(b) for a, b in ev2) / 1e3], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ts_, op=dist.ReduceOp.MAX)
            strong = {"value": gn * K / float(ts_[0]), "unit": UNIT, "global_envs": gn, "envs_per_gpu": ns, "ms_per_step": 1e3 * float(ts_[0]) / K,
                      "scaling": "strong", "what": "step only, the same 8192 environments split over the ranks"}
            del st2
        from brax_tracking_b200 import ppo
        tenv = presets.make_env(args.model, device=local)
        per_epoch = []
        tn = 8192
        ppo.train(tenv, num_timesteps=3 * tn * world * 16 * 32, episode_length=tenv.episode_length, num_envs=tn * world, num_evals=4,
                  learning_rate=3e-4, entropy_cost=1e-3, discounting=0.99, unroll_length=16, batch_size=tn * world, num_minibatches=32,
                  num_updates_per_batch=16, normalize_observations=True, run_evals=False,
                  progress_fn=lambda s_, mt: per_epoch.append(mt["training/device_sps"]))
        if rank == 0:
            train_blk = {"value": float(np.mean(per_epoch[1:])), "unit": UNIT, "envs_per_gpu": tn, "global_envs": tn * world, "scaling": "weak",
                         "training_steps_timed": len(per_epoch) - 1, "env_steps_per_training_step": tn * world * 16 * 32,
                         "what": "custom_ppo.train: rollout (policy inference + fused step, CUDA-graphed) + 16 x 32 minibatch updates "
                                 "(TF32 cuBLAS MLPs, fused tanh-Normal loss terms, one NCCL all-reduce of the flat gradient per minibatch, "
                                 "fused flat Adam); device time per epoch, max over ranks; train_fly.yaml hyper-parameters"}
    if rank == 0:
        total_env_steps = world * n * K
        value = total_env_steps / t_dev
        e2e = total_env_steps / t_e2e
        peak, peak_src = measured_peak_gbs()
        bytes_per_launch = ALGO_BYTES[args.model] * n
        avg_launch_s = float(np.mean(kernel_ms)) / 1e3
        achieved = bytes_per_launch / avg_launch_s / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": 1e3 * t_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.model, n), "envs_per_gpu": n, "global_envs": world * n,
                       "l2": "flushed between timed launches (256 MiB memset)", "parallelism": f"env-sharded x{world}, no collective in the step",
                       "warps_per_cta": env._native.warps_per_cta, "smem_bytes_per_cta": env._native.smem_bytes},
            "clocks": sampler.result(),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(n * m.nu * 4),
                    "d2h_bytes_per_step": int(n * (env.observation_size + 2) * 4), "wall_s": t_wall},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "bt_k_step", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(args.model), "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                         "avg_launch_ms": 1e3 * avg_launch_s},
        }
        if strong is not None:
            line["strong"] = strong
        if train_blk is not None:
            line["train"] = train_blk
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            # bounded sample of the same workload: ~10-30 CPU-seconds (wall time x threads) on the box's host cores
            ce = 128 * cores
            v, dt = cpu_port_run(args.model, ce, 8, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{ce} envs x 8 control steps, oracle port (C restatement of mjx.step, float32) + numpy env layer, "
                                              f"{dt:.1f} s wall = {dt * cores:.0f} CPU-s"}
        sys.stdout.flush()
        if saved_stdout is not None:
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=8192, help="envs per GPU")
    ap.add_argument("--model", default="rodent", choices=sorted(ALGO_BYTES))
    ap.add_argument("--ref-envs", type=int, default=None)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the strong-scaling and PPO-training blocks")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
