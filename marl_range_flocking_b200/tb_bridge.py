"""TensorBoard bridge for the batched rollout statistics.

The reference logs one set of scalars per game from its single env (`main.py:61-65`: Train/Buffer size,
Train/Sigma, Train/Score, Train/Average score, Train/Time taken; `learners/vdn/train_flock.py:111-113`:
train/score, train/epsilon, train/buffer_size). With E envs per GPU and W GPUs there is no "game": episodes
end all the time on the device. The batched equivalent reads the env's episode statistics (flushed by
reset on the device, `FLOCK_STAT_*`), all-reduces them over the ranks (`dist.allreduce_stats`: NCCL over
NVLink, the system's one collective) and writes, on rank 0 only, the same tags fed with the mean over the
episodes that closed since the previous call. One device->host read per call, none per step.
"""
from __future__ import annotations

import collections
import time
from typing import Dict, Optional

import torch

from . import dist as fdist


class StatsLogger:
    """`writer`: anything with `add_scalar(tag, value, step)` (e.g. `torch.utils.tensorboard.SummaryWriter`), or
    a log directory (a SummaryWriter is created on rank 0). `log()` returns the dict it wrote."""

    def __init__(self, writer, env, group=None, prefix: str = "Train", average_over: int = 10):
        self.env, self.group, self.prefix = env, group, prefix
        rank = torch.distributed.get_rank(group) if torch.distributed.is_initialized() else 0
        self.is_writer = rank == 0
        if isinstance(writer, str):
            if self.is_writer:
                from torch.utils.tensorboard import SummaryWriter
                writer = SummaryWriter(log_dir=writer)
            else:
                writer = None
        self.writer = writer
        self._last = torch.zeros(8, dtype=torch.int64)
        self._scores = collections.deque(maxlen=average_over)       # np.mean(scores[-10:]), main.py:64
        self._t0 = time.time()

    def log(self, step: int, extra: Optional[Dict[str, float]] = None) -> Dict[str, float]:
        """All-reduce the statistics, write the scalars of the episodes closed since the last call."""
        total = fdist.allreduce_stats(self.env.stats_tensor(), self.group).cpu()
        delta = total - self._last
        self._last = total
        n = max(int(delta[0]), 1)
        score = float(delta[2]) / 4294967296.0 / self.env.num_particles / n      # sum_t sum_i r / N, main.py:44
        if int(delta[0]) > 0:
            self._scores.append(score)
        out = {
            f"{self.prefix}/Score": score,
            f"{self.prefix}/Average score": sum(self._scores) / max(len(self._scores), 1),
            f"{self.prefix}/Episodes": float(total[0]),
            f"{self.prefix}/Mean episode length": float(delta[1]) / n,
            f"{self.prefix}/Reset attempts": float(delta[3]),
            f"{self.prefix}/Time taken": time.time() - self._t0,
        }
        for key, value in (extra or {}).items():          # e.g. Buffer size, Sigma, epsilon (main.py:61-62)
            out[f"{self.prefix}/{key}"] = float(value)
        if self.is_writer and self.writer is not None:
            for tag, value in out.items():
                self.writer.add_scalar(tag, value, step)
        self._t0 = time.time()
        return out
