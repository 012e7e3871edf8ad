"""Centralised Gymnasium-style environment for ONE wind farm (drop-in path).

API of the reference's ``WindFarmEnv`` (wfcrl/simple_env.py:13-99):

* ``reset(seed=None, options=None)`` returns the observation only (not ``(obs, info)``), like the reference;
* ``step(actions)`` takes ``{"yaw": array(T)}`` and returns ``(obs, reward[1], terminated=False, truncated, info)`` with
  ``info = {"power": MW (T,), "load": (T, 4)}``; the action array is modified in place when the actuation constraint
  freezes a turbine, as in the reference;
* attributes ``mdp``, ``action_space``, ``observation_space``, ``num_turbines``, ``max_num_steps``, ``controls``,
  ``num_moves``, ``accumulated_actions``, ``reward_shaper``, ``load_coef``, ``dt``, ``farm_case``.

The wake solve behind ``mdp.take_action`` runs on the GPU (``wfcrl_b200.interface.FlorisInterface``).  For thousands of
envs at once use ``wfcrl_b200.vector_env.VecWindFarmEnv``: the same semantics fused into one kernel launch per step.
"""
from __future__ import annotations

import copy
from typing import Dict

import numpy as np

from ._env_core import DUTY_LIMIT, busy_fraction, cooperative_reward
from .environments.data_cases import FarmCase
from .interface import BaseInterface
from .mdp import WindFarmMDP
from .rewards import DoNothingReward, RewardShaper

try:  # pragma: no cover - only where gymnasium is installed
    from gymnasium import Env as _EnvBase
except Exception:
    class _EnvBase:  # the two members of gymnasium.Env that callers of this class touch
        metadata: dict = {}

        @property
        def unwrapped(self):
            return self


class WindFarmEnv(_EnvBase):
    metadata = {"name": "centralized-windfarm"}

    def __init__(self, interface: BaseInterface, farm_case: FarmCase, controls: dict, continuous_control: bool = True,
                 reward_shaper: RewardShaper = None, start_iter: int = 0, max_num_steps: int = 500,
                 load_coef: float = 0.1):
        self.farm_case = farm_case
        self.controls = controls
        self.continuous_control = continuous_control
        self.max_num_steps = max_num_steps
        self.load_coef = load_coef
        self.dt = farm_case.dt
        self.reward_shaper = DoNothingReward() if reward_shaper is None else reward_shaper
        self.mdp = WindFarmMDP(interface, farm_case, controls, continuous_control=continuous_control,
                               start_iter=start_iter, horizon=start_iter + max_num_steps)
        self.num_turbines = self.mdp.num_turbines
        self.action_space = self.mdp.action_space
        self.observation_space = self.mdp.state_space
        self._begin_episode()

    # ------------------------------------------------------------------------------------------------------------
    def _begin_episode(self):
        self._state = self.mdp.start_state
        self.accumulated_actions = self.mdp.get_accumulated_actions()
        self.num_moves = 0

    def reset(self, seed=None, options=None):
        self.mdp.reset(seed, options)
        self.reward_shaper.reset()
        self._begin_episode()
        return copy.deepcopy(self._state)

    def step(self, actions: Dict):
        assert self._state is not None, "Call reset before `step`"
        self.num_moves += 1
        for name, command in actions.items():  # freeze actuators that exceeded their duty cycle (in place)
            rate = self.mdp.ACTUATORS_RATE.get(name)
            if rate is not None:
                busy = busy_fraction(self.accumulated_actions[name], rate, self.num_moves, self.farm_case.dt)
                command[busy >= DUTY_LIMIT] = 0.0

        state_before = self._state
        self._state, powers, loads, truncated = self.mdp.take_action(state_before, actions)
        raw = cooperative_reward(powers, loads, state_before["freewind_measurements"][0], self.load_coef)
        reward = np.array([self.reward_shaper(raw)])
        info = {"power": powers} if loads is None else {"power": powers, "load": loads}
        self.accumulated_actions = self.mdp.get_accumulated_actions()
        return copy.deepcopy(self._state), reward, False, truncated, info

    def close(self):
        pass
