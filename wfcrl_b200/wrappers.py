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


class VecLogWrapper:
    """History logger of a batched env (``make_vec(..., log=capacity)``): the reference's ``history`` keys
    (wfcrl/wrappers.py:24-52) as device-side ring buffers of the last ``capacity`` steps, ``[steps, B, ...]`` tensors,
    filled by device-to-device copies on the step's stream -- no host round trip per step."""

    def __init__(self, env, capacity: int = 1024):
        import torch

        if capacity <= 0:
            raise ValueError("capacity must be positive")
        self.env = env
        self.capacity = int(capacity)
        out = env.backend.out
        self._fields = {"yaw": "yaw", "wind_speed": "wind_speed", "wind_direction": "wind_direction",
                        "freewind_measurements": "freewind", "power": "power", "load": "load"}
        self._ring = {name: torch.zeros((self.capacity,) + tuple(out[src].shape), dtype=out[src].dtype, device=env.device)
                      for name, src in self._fields.items()}
        self._ring["reward"] = torch.zeros((self.capacity, env.num_envs), dtype=out["reward"].dtype, device=env.device)
        self._ring["truncated"] = torch.zeros((self.capacity, env.num_envs), dtype=torch.bool, device=env.device)
        self.num_logged = 0

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.__dict__["env"], name)

    def reset(self, seed=None, options=None, env_ids=None):
        if env_ids is None:
            self.num_logged = 0
        return self.env.reset(seed, options, env_ids)

    def step(self, action):
        result = self.env.step(action)
        _obs, reward, _terminated, truncated, _info = result
        info = self.env.last_info
        if isinstance(reward, dict):  # decentralised batch: every agent carries the shared reward / flag
            first = self.env.possible_agents[0]
            reward, truncated = reward[first], truncated[first]
        slot = self.num_logged % self.capacity
        out = self.env.backend.out
        final = info.get("final_observation")
        for name, src in self._fields.items():
            # on an auto-reset step log what the finished episode saw, not the restart rows
            if final is not None and name in final:
                self._ring[name][slot].copy_(final[name])
            elif final is not None and name in ("power", "load"):
                self._ring[name][slot].copy_(info["final_info"][name])
            else:
                self._ring[name][slot].copy_(out[src])
        self._ring["reward"][slot].copy_(reward)
        self._ring["truncated"][slot].copy_(truncated)
        self.num_logged += 1
        return result

    @property
    def history(self):
        """Chronological views of the logged steps: ``observation`` (dict), ``reward``, ``power``, ``load``, ``truncated``."""
        import torch

        n = min(self.num_logged, self.capacity)
        start = self.num_logged % self.capacity if self.num_logged > self.capacity else 0
        idx = (torch.arange(n, device=self.env.device) + start) % self.capacity

        def take(name):
            return self._ring[name][:n] if start == 0 else self._ring[name][idx]

        return {"observation": {k: take(k) for k in ("yaw", "freewind_measurements", "wind_speed", "wind_direction")},
                "reward": take("reward"), "power": take("power"), "load": take("load"), "truncated": take("truncated")}
