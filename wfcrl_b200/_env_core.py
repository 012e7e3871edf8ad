"""Arithmetic shared by the single-env Gymnasium and PettingZoo shells (host side of the drop-in path).

Both reference shells evaluate the same two expressions around ``WindFarmMDP.take_action``: the actuation-rate
constraint (wfcrl/simple_env.py:65-72, wfcrl/multiagent_env.py:198-207) and the cooperative reward
(wfcrl/simple_env.py:78-85, wfcrl/multiagent_env.py:219-227).  They live here once, with the reference's numpy dtypes
(float32 accumulators, float64 powers) so a batch of one reproduces the reference bit for bit; the batched path has the
same expressions fused into the step kernels.
"""
from __future__ import annotations

import numpy as np

DUTY_LIMIT = 0.1  # an actuator may be moving at most 10 % of the elapsed time


def busy_fraction(travel, rate, num_steps, dt):
    """Fraction of the elapsed time an actuator has been moving: accumulated travel / rate / steps / dt.

    ``travel`` is the float32 accumulator (array or scalar) of ``WindFarmMDP``; the three divisions are done one after
    the other exactly like the reference does, so the float32 rounding (and therefore the >= comparison) is identical."""
    moving_time = travel / rate
    return moving_time / num_steps / dt


def cooperative_reward(powers_mw, loads, reference_speed, load_coef):
    """Mean power in kW normalised by the cube of the free-stream speed of the state the action was taken in, minus the
    weighted mean absolute load proxy (when the simulator reports loads)."""
    reward = (powers_mw * 1e3 / (reference_speed ** 3)).mean()
    if loads is not None:
        reward = reward - load_coef * np.mean(np.abs(loads))
    return reward
