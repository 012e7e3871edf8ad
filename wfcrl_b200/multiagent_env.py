"""Decentralised PettingZoo-AEC environment for ONE wind farm (drop-in path): one agent per turbine.

API of the reference's ``MAWindFarmEnv`` (wfcrl/multiagent_env.py:15-257): agents ``turbine_1 .. turbine_T``; per-agent
dict spaces without the free-stream entry; ``step(action)`` acts for ``agent_selection``; the farm is advanced once per
cycle, when the last agent has acted, and every agent receives the same cooperative reward; the actuation constraint of
an agent uses its own snapshot of the accumulated travel, refreshed after each of its steps (so non-last agents lag one
joint action behind -- reproduced exactly).  Batched counterpart: ``wfcrl_b200.vector_env.VecMAWindFarmEnv``.
"""
from __future__ import annotations

import functools
from collections import OrderedDict

import numpy as np

from . import spaces
from ._env_core import DUTY_LIMIT, busy_fraction, cooperative_reward
from .aec import AECEnv, agent_selector
from .environments.data_cases import FarmCase
from .interface import BaseInterface
from .mdp import WindFarmMDP
from .rewards import DoNothingReward, RewardShaper

_GLOBAL_ONLY = "freewind_measurements"  # farm-level measurement, not part of an agent's local observation


class MAWindFarmEnv(AECEnv):
    metadata = {"name": "multiagent-windfarm", "is_parallelizable": True}

    def __init__(self, interface: BaseInterface, farm_case: FarmCase, controls: dict, continuous_control: bool = True,
                 reward_shaper: RewardShaper = None, start_iter: int = 0, max_num_steps: int = 500,
                 load_coef: float = 0.1):
        self.farm_case = farm_case
        self.controls = controls
        self.continuous_control = continuous_control
        self.max_num_steps = max_num_steps
        self.load_coef = load_coef
        self.reward_shaper = DoNothingReward() if reward_shaper is None else reward_shaper
        self.mdp = WindFarmMDP(interface, farm_case, controls, continuous_control=continuous_control,
                               start_iter=start_iter, horizon=start_iter + max_num_steps)
        self.num_turbines = self.mdp.num_turbines
        self.state_space = self.mdp.state_space
        self._state = None
        self.possible_agents = [f"turbine_{index}" for index in range(1, self.num_turbines + 1)]
        self.agent_name_mapping = {agent: index for index, agent in enumerate(self.possible_agents)}
        self._obs_spaces = {agent: self._local_boxes(self.mdp.state_space, index, skip=_GLOBAL_ONLY)
                            for agent, index in self.agent_name_mapping.items()}
        if continuous_control:
            self._action_spaces = {agent: self._local_boxes(self.mdp.action_space, index)
                                   for agent, index in self.agent_name_mapping.items()}
        else:
            self._action_spaces = {agent: {name: space[index] for name, space in self.mdp.action_space.items()}
                                   for agent, index in self.agent_name_mapping.items()}

    @staticmethod
    def _local_boxes(farm_space, index, skip=None):
        return {name: spaces.Box(box.low[index], box.high[index]) for name, box in farm_space.items() if name != skip}

    # -- spaces / observation ------------------------------------------------------------------------------------
    @functools.lru_cache(maxsize=None)
    def observation_space(self, agent):
        return self._obs_spaces[agent]

    @functools.lru_cache(maxsize=None)
    def action_space(self, agent):
        return self._action_spaces[agent]

    def state(self):
        return self._state

    def observe(self, agent):
        index = self.agent_name_mapping[agent]
        return OrderedDict((name, values[index]) for name, values in self._state.items() if name != _GLOBAL_ONLY)

    # -- AEC API --------------------------------------------------------------------------------------------------
    def reset(self, seed=None, options=None):
        self.mdp.reset(seed, options)
        self._state = self.mdp.start_state
        self.reward_shaper.reset()
        self.agents = list(self.possible_agents)
        travelled = self.mdp.get_accumulated_actions()
        self.accumulated_actions = {agent: {name: total[index] for name, total in travelled.items()}
                                    for agent, index in self.agent_name_mapping.items()}
        self._num_steps = dict.fromkeys(self.agents, 0)
        self.rewards = {agent: np.array([0.0]) for agent in self.agents}
        self._cumulative_rewards = {agent: np.array([0.0]) for agent in self.agents}
        self.terminations = dict.fromkeys(self.agents, False)
        self.truncations = dict.fromkeys(self.agents, False)
        self.infos = {agent: {} for agent in self.agents}
        self.actions = dict.fromkeys(self.agents)
        self.observations = {agent: self.observe(agent) for agent in self.agents}
        self.constrained = {agent: self.observe(agent) for agent in self.agents}
        self.num_moves = 0
        self._agent_selector = agent_selector(self.agents)
        self.agent_selection = self._agent_selector.next()

    def _validate(self, agent, action):
        unknown = [name for name in action if name not in self.mdp.controls]
        if unknown:
            raise ValueError(f"Control `{unknown[0]}` for agent {agent} is not activated."
                             f" List of activated controls: {list(self.mdp.controls.keys())}")
        if any(name not in action for name in self.mdp.controls):
            raise ValueError(f"Action {action} for agent {agent} is incomplete."
                             f" List of needed controls: {self.mdp.controls.keys()}")

    def _advance_farm(self):
        """All agents have spoken: one joint action, one wake solve, one shared reward."""
        joint = {name: np.zeros(self.num_turbines, dtype=np.float32) for name in self.mdp.controls}
        for index, action in enumerate(self.actions.values()):
            for name, command in action.items():
                joint[name][index] = np.asarray(command).reshape(-1)[0]
        state_before = self._state
        self._state, powers, loads, truncated = self.mdp.take_action(state_before, joint)
        raw = cooperative_reward(powers, loads, state_before["freewind_measurements"][0], self.load_coef)
        shared = np.array([self.reward_shaper(raw)])
        for agent, index in self.agent_name_mapping.items():
            if agent not in self.rewards:
                continue
            self.rewards[agent] = shared
            self.observations[agent] = self.observe(agent)
            self.truncations[agent] = truncated
            self.terminations[agent] = False
            self.infos[agent]["power"] = powers[index]
            if loads is not None:
                self.infos[agent]["load"] = loads[index]
        self.num_moves += 1

    def step(self, action):
        assert self._state is not None, "Call reset before `step`"
        agent = self.agent_selection
        if self.truncations[agent] or self.terminations[agent]:
            self._was_dead_step(action)
            return
        self._num_steps[agent] += 1
        self._validate(agent, action)

        snapshot = self.accumulated_actions[agent]
        for name, command in action.items():  # freeze this agent's actuators that exceeded their duty cycle (in place)
            rate = self.mdp.ACTUATORS_RATE.get(name)
            if rate is not None and busy_fraction(snapshot[name], rate, self._num_steps[agent],
                                                  self.farm_case.dt) >= DUTY_LIMIT:
                command[:] = 0.0

        self._cumulative_rewards[agent] = 0
        self.actions[agent] = action
        if self._agent_selector.is_last():
            self._advance_farm()
        else:
            self._clear_rewards()

        travelled = self.mdp.get_accumulated_actions()
        index = self.agent_name_mapping[agent]
        for name in action:
            snapshot[name] = travelled[name][index]
        self.agent_selection = self._agent_selector.next()
        self._accumulate_rewards()

    def close(self):
        pass
