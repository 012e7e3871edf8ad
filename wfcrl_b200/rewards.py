"""Reward shapers.

Drop-in for the reference's ``wfcrl.rewards`` (wfcrl/rewards.py:4-46): a shaper is called with the raw cooperative reward
of a step and returns the shaped one; ``reset()`` is invoked by the env at episode start.  The three shapers exist twice in
this package: here (host side, single-env path) and fused into the step kernels' epilogue for the batched path -- the
``kernel_code`` attribute is the link between the two (``FlorisBatch(reward_shaper=kernel_code)``).
"""
from __future__ import annotations

from abc import ABC


class RewardShaper(ABC):
    """Base class.  A subclass overrides ``__call__`` (the reference's abstract method, wfcrl/rewards.py:4-7) or ``shape``
    (what the shapers of this package implement); instances are callables either way."""

    #: name of the fused implementation in the step kernel, ``None`` for host-only shapers
    kernel_code = None

    def __call__(self, reward):
        return self.shape(reward)

    def shape(self, reward):
        raise NotImplementedError(f"{type(self).__name__} must override __call__ or shape")

    def update(self):
        """Hook kept for API compatibility (unused)."""

    def reset(self):
        """Called at every ``env.reset``; stateless shapers ignore it."""


class DoNothingReward(RewardShaper):
    """Pass the reward through unchanged."""

    kernel_code = "none"

    def shape(self, reward):
        return reward


class ReferencePercentage(RewardShaper):
    """Relative gain with respect to a constant ``reference`` reward."""

    kernel_code = "reference"

    def __init__(self, reference: float):
        self.reference = reference

    def shape(self, reward):
        gain = reward - self.reference
        return gain / self.reference


class StepPercentage(RewardShaper):
    """Relative gain with respect to the previous step's raw reward (0 until a non-zero reference exists)."""

    kernel_code = "step"

    def __init__(self, reference: float = 0.0):
        self.reference = reference

    def shape(self, reward):
        previous, self.reference = self.reference, reward
        if previous == 0:
            return 0.0
        return (reward - previous) / previous

    def reset(self, reference: float = 0.0):
        self.reference = reference
