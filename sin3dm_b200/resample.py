"""Timestep (schedule) samplers of the training loop — host numpy logic, same names and behaviour as the reference.

  reference: src/diffusion/resample.py:8-154 (``create_named_schedule_sampler``, ``UniformSampler``,
  ``LossSecondMomentResampler``); used by ``TrainLoop.forward_backward`` (train_util.py:209, 224-227).

Draws consume the global numpy RNG exactly like the reference (one ``np.random.choice(T, size=(B,), p=p)`` per call), so a seeded
run picks the same timesteps.  ``LossSecondMomentResampler`` works on current numpy (the reference's ``np.int`` alias is gone) and
synchronises ranks with one ``all_gather_object`` instead of three padded tensor all-gathers.
"""
import numpy as np
import torch as th
import torch.distributed as dist


class ScheduleSampler:
    """Importance sampler over diffusion steps: ``weights()`` -> positive array [T] (need not be normalised)."""

    def weights(self):
        raise NotImplementedError

    def sample(self, batch_size, device):
        """-> (timesteps int64 [B], loss weights fp32 [B] = 1 / (T * p[t])): the objective's mean is unchanged."""
        w = np.asarray(self.weights())
        p = w / np.sum(w)
        picked = np.random.choice(len(p), size=(batch_size,), p=p)
        scale = 1 / (len(p) * p[picked])
        return th.from_numpy(picked).long().to(device), th.from_numpy(scale).float().to(device)


class UniformSampler(ScheduleSampler):
    def __init__(self, diffusion):
        self.diffusion = diffusion
        self._weights = np.ones([diffusion.num_timesteps])

    def weights(self):
        return self._weights


class LossAwareSampler(ScheduleSampler):
    def update_with_local_losses(self, local_ts, local_losses):
        """Every rank contributes its (timestep, loss) pairs; all ranks then apply the same update in rank order."""
        mine = (local_ts.detach().cpu().tolist(), local_losses.detach().cpu().tolist())
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            everyone = [None] * dist.get_world_size()
            dist.all_gather_object(everyone, mine)
        else:
            everyone = [mine]
        ts = [int(t) for part in everyone for t in part[0]]
        losses = [float(v) for part in everyone for v in part[1]]
        self.update_with_all_losses(ts, losses)

    def update_with_all_losses(self, ts, losses):
        raise NotImplementedError


class LossSecondMomentResampler(LossAwareSampler):
    """Weights ~ sqrt(E[loss^2]) over the last ``history_per_term`` losses of each step, mixed with ``uniform_prob`` of uniform;
    uniform until every step has a full history."""

    def __init__(self, diffusion, history_per_term=10, uniform_prob=0.001):
        self.diffusion = diffusion
        self.history_per_term = history_per_term
        self.uniform_prob = uniform_prob
        T = diffusion.num_timesteps
        self._loss_history = np.zeros([T, history_per_term], dtype=np.float64)
        self._loss_counts = np.zeros([T], dtype=np.int64)

    def _warmed_up(self):
        return bool((self._loss_counts == self.history_per_term).all())

    def weights(self):
        T = self.diffusion.num_timesteps
        if not self._warmed_up():
            return np.ones([T], dtype=np.float64)
        rms = np.sqrt(np.mean(self._loss_history ** 2, axis=-1))
        return rms / np.sum(rms) * (1 - self.uniform_prob) + self.uniform_prob / T

    def update_with_all_losses(self, ts, losses):
        for t, loss in zip(ts, losses):
            n = self._loss_counts[t]
            if n < self.history_per_term:
                self._loss_history[t, n] = loss
                self._loss_counts[t] = n + 1
            else:                                       # full: drop the oldest entry
                self._loss_history[t] = np.append(self._loss_history[t, 1:], loss)


def create_named_schedule_sampler(name, diffusion):
    samplers = {"uniform": UniformSampler, "loss-second-moment": LossSecondMomentResampler}
    if name not in samplers:
        raise NotImplementedError(f"unknown schedule sampler: {name}")
    return samplers[name](diffusion)
