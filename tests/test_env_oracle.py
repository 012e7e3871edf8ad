"""CPU tests of the env-semantics oracle against the behavioural pins the reference's notebook holds
(spaces examples/demo.ipynb:98-99, 69 rows for max_num_steps=70 and the yaw trajectory :312-316)."""
import numpy as np

from oracle import c_oracle, env_oracle
from tests._util import layout


def _env(name="Ablaincourt_", **kw):
    lx, ly = layout(name)
    return env_oracle.EnvOracle(lx, ly, solver=c_oracle.solve, **kw)


def test_spaces_match_notebook_printout():
    env = _env()
    mdp = env.mdp
    assert np.all(mdp.action_low["yaw"] == np.float32(-5)) and np.all(mdp.action_high["yaw"] == np.float32(5))
    assert mdp.action_low["yaw"].shape == (7,) and mdp.action_low["yaw"].dtype == np.float32
    assert list(mdp.state_attributes) == ["yaw", "freewind_measurements", "wind_speed", "wind_direction"]
    assert np.all(mdp.low["yaw"] == -40) and np.all(mdp.high["yaw"] == 40)
    assert list(mdp.low["freewind_measurements"]) == [3.0, 0.0] and list(mdp.high["freewind_measurements"]) == [28.0, 360.0]
    assert np.all(mdp.low["wind_speed"] == 3) and np.all(mdp.high["wind_speed"] == 28)
    assert np.all(mdp.low["wind_direction"] == 0) and np.all(mdp.high["wind_direction"] == 360)


def test_notebook_episode_truncates_after_69_steps_with_the_stored_yaw_tail():
    """examples/demo.ipynb cell 13/16: max_num_steps=70 -> 69 history rows; yaws.tail() = -10 everywhere but T7=-5 at row 64."""
    env = _env(max_num_steps=70)
    env.reset(seed=0)
    history, i, done = [], 0, False
    while not done:
        joint = {"yaw": np.zeros(env.num_turbines)}
        if i % 5 == 0:
            joint["yaw"][int(i / 5 % env.num_turbines)] = -5.0
        obs, reward, term, trunc, info = env.step(joint)
        history.append(obs["yaw"].copy())
        i += 1
        done = term or trunc
    assert len(history) == 69
    tail = np.array(history[64:69])
    expect = np.full((5, 7), -10.0)
    expect[0, 6] = -5.0
    assert np.array_equal(tail, expect)
    assert reward.shape == (1,) and info["power"].shape == (7,) and info["load"].shape == (7, 4)


def test_reset_observation_is_the_kat1_vector():
    env = _env()
    obs = env.reset(options={"wind_speed": 6.48958384, "wind_direction": 266.363907})
    assert np.max(np.abs(obs["wind_speed"] - [6.46819497, 4.58929161, 6.46702757, 6.21243961, 6.20072934, 6.1100638,
                                              5.76785291])) < 2e-8
    assert np.allclose(obs["freewind_measurements"], [6.48958384, 266.363907])
    assert np.array_equal(obs["yaw"], np.zeros(7))


def test_actuation_constraint_zeroes_actions():
    """simple_env.py:65-72: acc / 0.3 / num_moves / dt >= 0.1  <=>  acc >= 1.8 * num_moves (float32 arithmetic)."""
    env = _env("Turb3_Row1_", max_num_steps=50)
    env.reset(options={"wind_speed": 8.0, "wind_direction": 270.0})
    yaws = []
    for _ in range(6):
        obs, *_ = env.step({"yaw": np.full(3, 5.0, dtype=np.float32)})
        yaws.append(float(obs["yaw"][0]))
    # step1: acc 0 -> +5 ; step2: acc 5 >= 3.6 -> blocked ; step3: 5 < 5.4 -> +5 ; step4: 10 >= 7.2 blocked ; ...
    assert yaws == [5.0, 5.0, 10.0, 10.0, 10.0, 15.0]


def test_reward_uses_previous_state_freestream_and_load_penalty():
    env = _env("Turb3_Row1_", load_coef=0.1)
    env.reset(options={"wind_speed": 2.0, "wind_direction": 270.0})  # clipped to 3.0 in the start state only
    _obs, r, _t, _tr, info = env.step({"yaw": np.zeros(3, dtype=np.float32)})
    expect = np.mean(info["power"] * 1e3 / 3.0 ** 3) - 0.1 * np.mean(np.abs(info["load"]))
    assert abs(r[0] - expect) < 1e-15
    _obs, r2, _t, _tr, info2 = env.step({"yaw": np.zeros(3, dtype=np.float32)})
    expect2 = np.mean(info2["power"] * 1e3 / 2.0 ** 3) - 0.1 * np.mean(np.abs(info2["load"]))
    assert abs(r2[0] - expect2) < 1e-15


def test_step_percentage_shaper():
    lx, ly = layout("Turb3_Row1_")
    env = env_oracle.EnvOracle(lx, ly, solver=c_oracle.solve, reward_shaper=env_oracle.StepPercentage())
    env.reset(options={"wind_speed": 8.0, "wind_direction": 270.0})
    _o, r1, *_ = env.step({"yaw": np.zeros(3, dtype=np.float32)})
    _o, r2, *_ = env.step({"yaw": np.array([5, 0, 0], dtype=np.float32)})
    assert r1[0] == 0.0 and r2[0] != 0.0


def test_multiagent_cycle_and_stale_constraint():
    """multiagent_env.py:198-249: non-last agents check the accumulator of two joint actions ago."""
    lx, ly = layout("Turb3_Row1_")
    env = env_oracle.MAEnvOracle(lx, ly, solver=c_oracle.solve, max_num_steps=20)
    env.reset(options={"wind_speed": 8.0, "wind_direction": 270.0})
    yaw_hist = []
    for cycle in range(5):
        for agent in list(env.agents):
            assert env.agent_selection == agent
            env.step({"yaw": np.array([5.0])})
        yaw_hist.append(env._state["yaw"].copy())
    yaw_hist = np.array(yaw_hist)
    # last agent (fresh accumulator) behaves like the centralized env: 5, 5, 10, 10, 10
    assert list(yaw_hist[:, 2]) == [5.0, 5.0, 10.0, 10.0, 10.0]
    # non-last agents see a one-cycle-stale accumulator (A_{n-2}): cycle 2 still sees 0, cycle 3 sees 5 < 5.4
    assert list(yaw_hist[:, 0]) == [5.0, 10.0, 15.0, 15.0, 15.0]
    obs, rew, term, trunc, info = env.last()
    assert set(obs.keys()) == {"yaw", "wind_speed", "wind_direction"} and not term
    assert "power" in info and "load" in info


def test_multiagent_truncation_and_agent_iter_terminates():
    lx, ly = layout("Turb3_Row1_")
    env = env_oracle.MAEnvOracle(lx, ly, solver=c_oracle.solve, max_num_steps=5)
    env.reset(options={"wind_speed": 8.0, "wind_direction": 270.0})
    n_live = 0
    for agent in env.agent_iter():
        obs, r, term, trunc, info = env.last()
        if term or trunc:
            env.step(None)
        else:
            env.step({"yaw": np.zeros(1)})
            n_live += 1
    assert n_live == 3 * 4  # max_num_steps=5 -> 4 cycles before truncation
    assert env.agents == []


def test_philox_known_answers_and_reset_distribution():
    """Checker of the library's reset sampler: Philox4x32-10 against the Random123 known-answer vectors, and the sampled
    (wind speed, wind direction) against the reference's reset distribution (mdp.py:242-258) drawn with numpy."""
    from scipy import stats

    assert env_oracle.philox4x32_10((0, 0, 0, 0), (0, 0)) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert env_oracle.philox4x32_10((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert env_oracle.philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    n = 3000
    draws = np.array([env_oracle.sampled_reset_wind(11, g, e) for g in range(n // 2) for e in range(2)])
    rng = np.random.default_rng(5)
    ref_ws = np.clip(8 * rng.weibull(8, 20000), 3, 28)
    ref_wd = np.clip(rng.normal(270, 20, 20000) % 360, 0, 360)
    assert stats.ks_2samp(draws[:, 0], ref_ws).pvalue > 1e-3
    assert stats.ks_2samp(draws[:, 1], ref_wd).pvalue > 1e-3
    assert abs(np.corrcoef(draws[:, 0], draws[:, 1])[0, 1]) < 0.06       # independent draws
    assert abs(np.corrcoef(draws[0::2, 0], draws[1::2, 0])[0, 1]) < 0.08  # consecutive episodes of an env
    ws, wd, ti = env_oracle.sampled_reset_wind(11, 3, 0, ti_range=(0.04, 0.12))
    assert (ws, wd) == tuple(draws[6]) and 0.04 <= ti < 0.12
