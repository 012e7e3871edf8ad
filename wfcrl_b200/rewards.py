"""Reward shapers with the call protocol of the reference (wfcrl/rewards.py:4-46): ``shaper(reward)`` and ``reset()``.

In the batched path the same three shapers are fused into the kernel epilogue (``WfConfig.reward_shaper``); these
host-side classes serve the single-env drop-in path and carry the ``kernel_code`` the batched path needs."""
from __future__ import annotations

from abc import ABC, abstractmethod


class RewardShaper(ABC):
    kernel_code = None  # name understood by FlorisBatch(reward_shaper=...); None = host-only shaper

    @abstractmethod
    def __call__(self, reward: float):
        ...

    def update(self):
        pass

    def reset(self):
        pass


class DoNothingReward(RewardShaper):
    """Identity."""

    kernel_code = "none"

    def __call__(self, reward):
        return reward


class ReferencePercentage(RewardShaper):
    """Relative improvement over a fixed reference."""

    kernel_code = "reference"

    def __init__(self, reference: float):
        self.reference = reference

    def __call__(self, reward):
        return (reward - self.reference) / self.reference


class StepPercentage(RewardShaper):
    """Relative improvement over the previous step's reward; 0 while the reference is 0."""

    kernel_code = "step"

    def __init__(self, reference: float = 0.0):
        self.reference = reference

    def __call__(self, reward):
        shaped = 0.0 if self.reference == 0 else (reward - self.reference) / self.reference
        self.reference = reward
        return shaped

    def reset(self, reference: float = 0.0):
        self.reference = reference
