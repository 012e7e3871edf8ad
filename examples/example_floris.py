"""Decentralised single-env loop on the B200 backend: same flow as the reference's examples/example_floris.py
(PettingZoo AEC iteration, StepPercentage reward shaping), only the import changes."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wfcrl_b200 import environments as envs  # noqa: E402
from wfcrl_b200.rewards import StepPercentage  # noqa: E402

env = envs.make("Dec_Ablaincourt_Floris", max_num_steps=100, reward_shaper=StepPercentage(), load_coef=1)


def policy(agent, step):
    if agent == "turbine_1" and step == 20:
        return {"yaw": np.array([15.0])}
    return {"yaw": np.array([0.0])}


env.reset()
totals = {agent: 0 for agent in env.possible_agents}
done = {agent: False for agent in env.possible_agents}
steps = {agent: 0 for agent in env.possible_agents}
for agent in env.agent_iter():
    observation, reward, termination, truncation, info = env.last()
    done[agent] = done[agent] or termination or truncation
    totals[agent] += reward
    if done[agent]:
        action = None
    else:
        action = policy(agent, steps[agent])
        steps[agent] += 1
    env.step(action)
print(f"Total reward = {totals}")
