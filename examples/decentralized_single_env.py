"""One decentralised wind farm (PettingZoo AEC API) on the B200 backend.

Counterpart of the reference's ``examples/example_floris.py``: Ablaincourt, one agent per turbine, StepPercentage reward
shaping, a scripted policy that yaws the first turbine once.  Only the import changes with respect to ifpen/wfcrl-env.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wfcrl_b200 import environments as envs  # noqa: E402
from wfcrl_b200.rewards import StepPercentage  # noqa: E402

KICK_AGENT, KICK_STEP, KICK_DEG = "turbine_1", 20, 15.0


def scripted_policy(agent: str, step: int):
    """Hold every yaw, except a single 15 degree request on the first turbine (clipped to the 5 degree step by the env)."""
    return {"yaw": np.array([KICK_DEG if (agent == KICK_AGENT and step == KICK_STEP) else 0.0])}


def run_episode(env):
    env.reset()
    returns = dict.fromkeys(env.possible_agents, 0.0)
    acted = dict.fromkeys(env.possible_agents, 0)
    finished = set()
    for agent in env.agent_iter():
        _obs, reward, terminated, truncated, _info = env.last()
        returns[agent] += float(np.asarray(reward).reshape(-1)[0])
        if terminated or truncated:
            finished.add(agent)
        if agent in finished:
            env.step(None)
            continue
        env.step(scripted_policy(agent, acted[agent]))
        acted[agent] += 1
    return returns


if __name__ == "__main__":
    farm = envs.make("Dec_Ablaincourt_Floris", max_num_steps=100, reward_shaper=StepPercentage(), load_coef=1)
    print("Total reward =", run_episode(farm))
