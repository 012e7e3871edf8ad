"""Minimal PettingZoo-AEC base used when pettingzoo is not installed.

Implements the bookkeeping of ``pettingzoo.AECEnv`` 1.24.3 that ``MAWindFarmEnv`` relies on:
``last()``, ``agent_iter()``, ``_was_dead_step``, ``_clear_rewards``, ``_accumulate_rewards`` and ``agent_selector``.
"""
from __future__ import annotations

try:  # pragma: no cover
    from pettingzoo import AECEnv  # noqa: F401
    from pettingzoo.utils import agent_selector  # noqa: F401

    HAVE_PETTINGZOO = True
except Exception:
    HAVE_PETTINGZOO = False

    class agent_selector:  # noqa: N801 - name of the pettingzoo utility
        def __init__(self, agent_order):
            self.reinit(agent_order)

        def reinit(self, agent_order):
            self.agent_order = list(agent_order)
            self._current_agent = 0
            self.selected_agent = 0

        def reset(self):
            self.reinit(self.agent_order)
            return self.next()

        def next(self):
            self._current_agent = (self._current_agent + 1) % len(self.agent_order)
            self.selected_agent = self.agent_order[self._current_agent - 1]
            return self.selected_agent

        def is_last(self):
            return self.selected_agent == self.agent_order[-1]

        def is_first(self):
            return self.selected_agent == self.agent_order[0]

    class AECEnv:
        metadata = {}

        @property
        def unwrapped(self):
            return self

        @property
        def num_agents(self):
            return len(self.agents)

        @property
        def max_num_agents(self):
            return len(self.possible_agents)

        def observe(self, agent):
            raise NotImplementedError

        def last(self, observe: bool = True):
            agent = self.agent_selection
            observation = self.observe(agent) if observe else None
            return (observation, self._cumulative_rewards[agent], self.terminations[agent], self.truncations[agent],
                    self.infos[agent])

        def agent_iter(self, max_iter: int = 2 ** 63):
            count = 0
            while self.agents and count < max_iter:
                count += 1
                yield self.agent_selection

        def _clear_rewards(self):
            for agent in self.rewards:
                self.rewards[agent] = 0

        def _accumulate_rewards(self):
            for agent, reward in self.rewards.items():
                self._cumulative_rewards[agent] += reward

        def _deads_step_first(self):
            dead = [a for a in self.agents if self.terminations[a] or self.truncations[a]]
            if dead:
                self._skip_agent_selection = self.agent_selection
                self.agent_selection = dead[0]
            return self.agent_selection

        def _was_dead_step(self, action):
            if action is not None:
                raise ValueError("when an agent is dead, the only valid action is None")
            agent = self.agent_selection
            assert self.terminations[agent] or self.truncations[agent]
            for store in (self.rewards, self._cumulative_rewards, self.terminations, self.truncations, self.infos):
                del store[agent]
            self.agents.remove(agent)
            dead = [a for a in self.agents if self.terminations[a] or self.truncations[a]]
            if dead:
                if getattr(self, "_skip_agent_selection", None) is None:
                    self._skip_agent_selection = self.agent_selection
                self.agent_selection = dead[0]
            else:
                if getattr(self, "_skip_agent_selection", None) is not None:
                    self.agent_selection = self._skip_agent_selection
                self._skip_agent_selection = None
            self._clear_rewards()

        def close(self):
            pass
