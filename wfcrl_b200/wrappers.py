"""History-logging wrappers with the attributes of the reference's ``LogWrapper`` / ``AECLogWrapper``
(wfcrl/wrappers.py:24-88): ``history`` lists of observation / reward / load / power, attribute passthrough."""
from __future__ import annotations


class _Passthrough:
    def __init__(self, env):
        self.env = env
        self.continuous_control = env.continuous_control
        self.max_num_steps = env.max_num_steps
        self._state = env.mdp.start_state
        self.num_turbines = env.mdp.num_turbines
        self.mdp = env.mdp
        self.controls = env.controls

    def __getattr__(self, name):
        if name.startswith("_") and name != "_state":
            raise AttributeError(name)
        return getattr(self.__dict__["env"], name)

    @property
    def unwrapped(self):
        return getattr(self.env, "unwrapped", self.env)


def _empty_history():
    return {"observation": [], "reward": [], "load": [], "power": []}


class LogWrapper(_Passthrough):
    def __init__(self, env):
        super().__init__(env)
        self.history = _empty_history()

    def step(self, action):
        observation, reward, terminated, truncated, info = self.env.step(action)
        self.history["observation"].append(observation)
        self.history["reward"].append(reward)
        for key in ("power", "load"):
            if key in info:
                self.history[key].append(info[key])
        return observation, reward, terminated, truncated, info

    def reset(self, seed=None, options=None):
        self.history = _empty_history()
        return self.env.reset(seed, options)


class AECLogWrapper(_Passthrough):
    def __init__(self, env):
        super().__init__(env)
        self.history = {agent: _empty_history() for agent in env.possible_agents}

    def last(self):
        agent = self.env.agent_selection
        observation, reward, termination, truncation, info = self.env.last()
        log = self.history[agent]
        log["observation"].append(observation)
        log["reward"].append(reward)
        for key in ("power", "load"):
            if key in info:
                log[key].append(info[key])
        return observation, reward, termination, truncation, info

    def step(self, action):
        return self.env.step(action)

    def agent_iter(self, max_iter: int = 2 ** 63):
        return self.env.agent_iter(max_iter)

    def reset(self, seed=None, options=None):
        self.history = {agent: _empty_history() for agent in self.env.possible_agents}
        return self.env.reset(seed, options)


class RandomSimulator(_Passthrough):
    """Calls the interface's (stub) ``sample_parameters`` on every reset (wfcrl/wrappers.py:6-21)."""

    def __init__(self, env):
        super().__init__(env)
        self.parameters_vector = env.mdp.interface.get_parameters()

    def reset(self, seed=None, options=None):
        self.parameters_vector = self.env.mdp.interface.sample_parameters()
        return self.env.reset(seed, options)
