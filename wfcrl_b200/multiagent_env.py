"""Decentralised (one agent per turbine) PettingZoo-AEC environment, single env, drop-in path.

Public surface of the reference's ``MAWindFarmEnv`` (wfcrl/multiagent_env.py:15-257): agents ``turbine_1..T``, per-agent
dict spaces without the free-stream entry, AEC ``step(action)`` for ``agent_selection`` with the wake solve executed
once per cycle when the last agent has acted, the same cooperative reward for every agent, and the per-agent actuation
constraint with its one-cycle-stale accumulator snapshot for non-last agents.
The batched counterpart is ``wfcrl_b200.vector_env.VecMAWindFarmEnv`` (parallel-API style, whole cycles per call).
"""
from __future__ import annotations

import functools
from collections import OrderedDict

import numpy as np

from . import spaces
from .aec import AECEnv, agent_selector
from .environments.data_cases import FarmCase
from .interface import BaseInterface
from .mdp import WindFarmMDP
from .rewards import DoNothingReward, RewardShaper


class MAWindFarmEnv(AECEnv):
    metadata = {"name": "multiagent-windfarm", "is_parallelizable": True}

    def __init__(self, interface: BaseInterface, farm_case: FarmCase, controls: dict, continuous_control: bool = True,
                 reward_shaper: RewardShaper = None, start_iter: int = 0, max_num_steps: int = 500,
                 load_coef: float = 0.1):
        self.mdp = WindFarmMDP(interface=interface, farm_case=farm_case, controls=controls,
                               continuous_control=continuous_control, start_iter=start_iter,
                               horizon=start_iter + max_num_steps)
        self.continuous_control = continuous_control
        self.max_num_steps = max_num_steps
        self._state = None
        self.num_turbines = self.mdp.num_turbines
        self.reward_shaper = reward_shaper if reward_shaper is not None else DoNothingReward()
        self.controls = controls
        self.farm_case = farm_case
        self.state_space = self.mdp.state_space
        self.load_coef = load_coef
        self.possible_agents = [f"turbine_{k + 1}" for k in range(self.num_turbines)]
        self.agent_name_mapping = {name: k for k, name in enumerate(self.possible_agents)}
        self._build_agent_spaces()

    # -- spaces ---------------------------------------------------------------------------------------------------
    def _build_agent_spaces(self):
        self._obs_spaces, self._action_spaces = {}, {}
        for k, agent in enumerate(self.possible_agents):
            self._obs_spaces[agent] = {
                key: spaces.Box(box.low[k], box.high[k]) for key, box in self.mdp.state_space.items()
                if key != "freewind_measurements"}
            if self.continuous_control:
                self._action_spaces[agent] = {
                    key: spaces.Box(box.low[k], box.high[k]) for key, box in self.mdp.action_space.items()}
            else:
                self._action_spaces[agent] = {key: space[k] for key, space in self.mdp.action_space.items()}

    @functools.lru_cache(maxsize=None)
    def observation_space(self, agent):
        return self._obs_spaces[agent]

    @functools.lru_cache(maxsize=None)
    def action_space(self, agent):
        return self._action_spaces[agent]

    def state(self):
        return self._state

    def observe(self, agent):
        k = self.agent_name_mapping[agent]
        return OrderedDict((key, values[k]) for key, values in self.state().items()
                           if key != "freewind_measurements")

    def _join_actions(self, agent_actions):
        joint = {control: np.zeros(self.num_turbines, dtype=np.float32) for control in self.mdp.controls}
        for k, action in enumerate(agent_actions.values()):
            for control in action:
                joint[control][k] = np.asarray(action[control]).reshape(-1)[0]
        return joint

    # -- AEC API --------------------------------------------------------------------------------------------------
    def reset(self, seed=None, options=None):
        self.mdp.reset(seed, options)
        self._state = self.mdp.start_state
        self.reward_shaper.reset()
        self.agents = self.possible_agents[:]
        self._num_steps = {agent: 0 for agent in self.agents}
        self.rewards = {agent: np.array([0.0]) for agent in self.agents}
        self._cumulative_rewards = {agent: np.array([0.0]) for agent in self.agents}
        self.terminations = {agent: False for agent in self.agents}
        self.truncations = {agent: False for agent in self.agents}
        self.infos = {agent: {} for agent in self.agents}
        self.actions = {agent: None for agent in self.agents}
        self.observations = {agent: self.observe(agent) for agent in self.agents}
        self.constrained = {agent: self.observe(agent) for agent in self.agents}
        totals = self.mdp.get_accumulated_actions()
        self.accumulated_actions = {agent: {control: totals[control][k] for control in totals}
                                    for k, agent in enumerate(self.agents)}
        self.num_moves = 0
        self._agent_selector = agent_selector(self.agents)
        self.agent_selection = self._agent_selector.next()

    def step(self, action):
        assert self._state is not None, "Call reset before `step`"
        agent = self.agent_selection
        if self.truncations[agent] or self.terminations[agent]:
            self._was_dead_step(action)
            return
        self._num_steps[agent] += 1
        for control in action:
            if control not in self.mdp.controls:
                raise ValueError(f"Control `{control}` for agent {agent} is not activated."
                                 f" List of activated controls: {list(self.mdp.controls.keys())}")
        if any(control not in action for control in self.mdp.controls):
            raise ValueError(f"Action {action} for agent {agent} is incomplete."
                             f" List of needed controls: {self.mdp.controls.keys()}")

        # actuation constraint on this agent's (snapshot of the) accumulated travel
        snapshot = self.accumulated_actions[agent]
        for control in action:
            rate = self.mdp.ACTUATORS_RATE.get(control)
            if rate is None:
                continue
            busy_frac = snapshot[control] / rate / self._num_steps[agent] / self.farm_case.dt
            if busy_frac >= 0.1:
                action[control][:] = 0.0

        self._cumulative_rewards[agent] = 0
        self.actions[agent] = action

        if self._agent_selector.is_last():  # the whole farm moves once every agent has spoken
            previous_state = self.state()
            next_state, powers, loads, truncated = self.mdp.take_action(self._state, self._join_actions(self.actions))
            reward = (powers * 1e3 / (previous_state["freewind_measurements"][0] ** 3)).mean()
            if loads is not None:
                reward = reward - self.load_coef * np.mean(np.abs(loads))
            reward = np.array([self.reward_shaper(reward)])
            self._state = next_state
            for name in self.agents:
                k = self.agent_name_mapping[name]
                if loads is not None:
                    self.infos[name]["load"] = loads[k]
                self.rewards[name] = reward
                self.observations[name] = self.observe(name)
                self.truncations[name] = truncated
                self.terminations[name] = False
                self.infos[name]["power"] = powers[k]
            self.num_moves += 1
        else:
            self._clear_rewards()

        totals = self.mdp.get_accumulated_actions()
        for control in action:
            self.accumulated_actions[agent][control] = totals[control][self.agent_name_mapping[agent]]
        self.agent_selection = self._agent_selector.next()
        self._accumulate_rewards()

    def close(self):
        pass
