#!/usr/bin/env python
"""Full PPO imitation training on the fused B200 step (BASELINE.json configs[4]); the torch counterpart of the
reference's `python main.py` (/root/reference/main.py:48-334) without the Hydra / wandb / rendering glue.

    python train.py --model rodent --num-envs 8192 --num-timesteps 20000000
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 train.py --num-envs 65536 ...

Hyper-parameters default to /root/reference/configs/train/train_fly.yaml (batch = num_envs, 32 minibatches, 16 updates
per batch, unroll 16, lr 3e-4, entropy 1e-3, gamma 0.99, clip 0.3)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from brax_tracking_b200 import ppo, presets  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="rodent", choices=["rodent", "fly_free", "fly_tethered", "rodent_pair"])
    ap.add_argument("--num-envs", type=int, default=8192)
    ap.add_argument("--num-timesteps", type=int, default=10_000_000)
    ap.add_argument("--num-evals", type=int, default=5)
    ap.add_argument("--batch-size", type=int, default=None)
    ap.add_argument("--num-minibatches", type=int, default=32)
    ap.add_argument("--num-updates-per-batch", type=int, default=16)
    ap.add_argument("--unroll-length", type=int, default=16)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--checkpoint", default=None, help="resume from this file (written by --checkpoint-dir)")
    ap.add_argument("--checkpoint-dir", default=None, help="rank 0 saves <dir>/<env_steps>.pt after every evaluation and at the end")
    ap.add_argument("--num-eval-envs", type=int, default=128)
    ap.add_argument("--deterministic-eval", action="store_true")
    ap.add_argument("--no-eval", action="store_true")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries the progress lines only: NCCL prints its version banner to fd 1 when the communicator is created
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved, 1)
    env = presets.make_env(a.model, device=local)

    def progress(step, metrics):
        print(json.dumps({"env_steps": step, **{k: round(v, 5) for k, v in metrics.items()}}), flush=True)

    ppo.train(env, num_timesteps=a.num_timesteps, episode_length=env.episode_length, num_envs=a.num_envs, num_evals=a.num_evals,
              learning_rate=3e-4, entropy_cost=1e-3, discounting=0.99, seed=a.seed, unroll_length=a.unroll_length,
              batch_size=a.batch_size or a.num_envs, num_minibatches=a.num_minibatches, num_updates_per_batch=a.num_updates_per_batch,
              normalize_observations=True, reward_scaling=1.0, progress_fn=progress, restore_checkpoint_path=a.checkpoint,
              checkpoint_dir=a.checkpoint_dir, num_eval_envs=a.num_eval_envs, deterministic_eval=a.deterministic_eval,
              run_evals=not a.no_eval)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
