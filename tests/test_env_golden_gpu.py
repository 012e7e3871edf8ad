"""The CUDA paths against the step records of the UNMODIFIED reference env classes (tests/golden/env_ref_*.json, see
tools/make_golden_env.py): (1) the single-env drop-ins behind ``envs.make`` (FP64 kernels through FlorisInterface),
(2) the batched vector envs (fused step kernel, FP64 and FP32).  Integers, booleans and the float32 yaw state must be
bit-exact; floats within 1e-9 (FP64) / 1e-4 (FP32, loads 2e-4 with an absolute floor of 1e-5 m/s)."""
import numpy as np
import pytest

from tests import _golden_env as G

pytestmark = pytest.mark.gpu


def _product_shaper(mk):
    from wfcrl_b200 import rewards

    kind, ref = G.shaper_spec(mk)
    return {"none": None, "reference": rewards.ReferencePercentage(ref) if kind == "reference" else None,
            "step": rewards.StepPercentage(ref) if kind == "step" else None}[kind]


def _make_kwargs(mk):
    kw = {k: mk[k] for k in ("max_num_steps", "load_coef", "continuous_control", "log") if k in mk}
    if "controls" in mk:
        kw["controls"] = {k: tuple(v) for k, v in mk["controls"].items()}
    shaper = _product_shaper(mk)
    if shaper is not None:
        kw["reward_shaper"] = shaper
    if "wind_time_series" in mk:
        kw["wind_time_series"] = np.array(mk["wind_time_series"])
    return kw


@pytest.mark.parametrize("name", G.SINGLE)
def test_drop_in_single_env_reproduces_reference(cuda_device, name):
    from wfcrl_b200 import environments as envs

    scen = G.load(name)
    if scen.get("np_seed") is not None:
        np.random.seed(scen["np_seed"])
    env = envs.make(scen["env_id"], **_make_kwargs(scen["make_kwargs"]))
    rk = scen["reset_kwargs"]
    obs = env.reset(seed=rk.get("seed"), options=rk.get("options"))
    rec = scen["record"]
    G.compare_obs(obs, rec["reset_observation"], 1e-9, "reset")
    for k, st in enumerate(rec["steps"]):
        obs, reward, term, trunc, info = env.step({"yaw": G.arr(st["action"]).copy()})
        G.compare_obs(obs, st["observation"], 1e-9, k)
        assert term is False and bool(trunc) == st["truncated"], k
        assert G.close(reward, G.arr(st["reward"]), 1e-9), (k, reward, st["reward"])
        assert G.close(info["power"], G.arr(st["power"]), 1e-9)
        assert G.close(info["load"], G.arr(st["load"]), 1e-8, 1e-9)
    if "history_lengths" in rec:
        assert {k: len(v) for k, v in env.history.items()} == rec["history_lengths"]


@pytest.mark.parametrize("name", G.MULTI)
def test_drop_in_aec_env_reproduces_reference(cuda_device, name):
    from wfcrl_b200 import environments as envs

    scen = G.load(name)
    env = envs.make(scen["env_id"], **_make_kwargs(scen["make_kwargs"]))
    rk = scen["reset_kwargs"]
    env.reset(seed=rk.get("seed"), options=rk.get("options"))
    rec = scen["record"]
    assert list(env.possible_agents) == rec["agents"]
    events = iter(rec["events"])
    n = 0
    for agent in env.agent_iter():
        ev = next(events)
        assert agent == ev["agent"], n
        obs, reward, term, trunc, info = env.last()
        G.compare_obs(obs, ev["observation"], 1e-9, n)
        assert G.close(reward, G.arr(ev["cumulative_reward"]), 1e-9), (n, reward)
        assert bool(term) == ev["terminated"] and bool(trunc) == ev["truncated"], n
        assert set(info) == set(ev["info"]), n
        for key, val in ev["info"].items():
            assert G.close(info[key], G.arr(val), 1e-8, 1e-9), (n, key)
        env.step(None if ev["action"] is None else {"yaw": G.arr(ev["action"]).copy()})
        n += 1
    assert next(events, None) is None and list(env.agents) == rec["agents_left"]


def _vec_env(scen, precision, multi_agent=False):
    from wfcrl_b200 import environments as envs

    mk = dict(scen["make_kwargs"])
    mk.pop("log", None)
    kw = _make_kwargs(mk)
    # three copies of the same env in one batch: every row must reproduce the reference record
    return envs.make_vec(scen["env_id"], 3, precision=precision, auto_reset=False, **kw)


@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("name", [n for n in G.SINGLE if "time_series" not in n])
def test_batched_env_reproduces_reference(cuda_device, name, precision):
    import torch

    scen = G.load(name)
    env = _vec_env(scen, precision)
    rk = scen["reset_kwargs"]
    rec = scen["record"]
    if "seed" in rk:  # the batched reset draws env g's wind as the reference's reset(seed + g): row 0 is the record's env
        obs = env.reset(seed=rk["seed"])
        rows = [0]
    else:
        obs = env.reset(options=rk.get("options"))
        rows = [0, 1, 2]
    tol = 1e-9 if precision == "f64" else 1e-4
    ltol, lfloor = (1e-8, 1e-9) if precision == "f64" else (2e-4, 5e-2)  # loads: |d| <= ltol * max(|ref|, floor)

    def check_obs(obs, ref, where):
        for b in rows:
            got = {"yaw": obs["yaw"][b].cpu().numpy(), "freewind_measurements": obs["freewind_measurements"][b].cpu().numpy(),
                   "wind_speed": obs["wind_speed"][b].cpu().numpy(), "wind_direction": obs["wind_direction"][b].cpu().numpy()}
            G.compare_obs(got, ref, tol, where)

    check_obs(obs, rec["reset_observation"], "reset")
    for k, st in enumerate(rec["steps"]):
        a = torch.as_tensor(np.tile(G.arr(st["action"]).astype(np.float32), (3, 1)), device="cuda")
        obs, reward, term, trunc, info = env.step(a)
        torch.cuda.synchronize()
        check_obs(obs, st["observation"], k)
        for b in rows:
            assert not bool(term[b]) and bool(trunc[b]) == st["truncated"], (k, b)
            assert G.close(reward[b].cpu().numpy(), G.arr(st["reward"])[0], tol), (k, b, reward[b], st["reward"])
            assert G.close(info["power"][b].cpu().numpy(), G.arr(st["power"]), tol, 1e-6), (k, b)
            assert G.close(info["load"][b].cpu().numpy(), G.arr(st["load"]), ltol, lfloor), (k, b)
    env.close()


@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("name", G.MULTI)
def test_batched_multi_agent_env_reproduces_reference(cuda_device, name, precision):
    """One VecMAWindFarmEnv.step = one full agent cycle of the reference's AEC env: compare at every cycle boundary what the
    first agent sees through last() (observation, reward of the cycle, truncation, power / load)."""
    import torch

    scen = G.load(name)
    env = _vec_env(scen, precision, multi_agent=True)
    rk = scen["reset_kwargs"]
    rec = scen["record"]
    agents = rec["agents"]
    T = len(agents)
    if "seed" in rk:
        env.reset(seed=rk["seed"])
    else:
        env.reset(options=rk.get("options"))
    tol = 1e-9 if precision == "f64" else 1e-4
    live = [ev for ev in rec["events"] if not ev["dead_step"]]
    n_cycles = len(live) // T
    for c in range(n_cycles):
        cyc = live[c * T:(c + 1) * T]
        assert [ev["agent"] for ev in cyc] == agents
        joint = np.array([G.arr(ev["action"])[0] for ev in cyc], dtype=np.float32)
        obs, reward, term, trunc, info = env.step(torch.as_tensor(np.tile(joint, (3, 1)), device="cuda"))
        torch.cuda.synchronize()
        # what every agent sees at its NEXT last(): the state after this joint step
        nxt = rec["events"][(c + 1) * T:(c + 2) * T]
        for j, ev in enumerate(nxt):
            a = agents[j]
            assert ev["agent"] == a
            want = {k: G.arr(v) for k, v in ev["observation"].items()}
            assert np.array_equal(obs[a]["yaw"][0].cpu().numpy().astype(np.float64), want["yaw"].astype(np.float64)), (c, a)
            assert G.close(obs[a]["wind_speed"][0].cpu().numpy(), want["wind_speed"], tol), (c, a)
            assert G.close(obs[a]["wind_direction"][0].cpu().numpy(), want["wind_direction"], tol), (c, a)
            assert bool(trunc[a][0]) == ev["truncated"], (c, a)
            assert G.close(info[a]["power"][0].cpu().numpy(), G.arr(ev["info"]["power"]), tol, 1e-6), (c, a)
        # the last agent of the cycle collects the cycle's reward un-accumulated (multiagent_env.py:247-249)
        assert G.close(reward[agents[-1]][0].cpu().numpy(), G.arr(nxt[-1]["cumulative_reward"])[0], tol), c
    env.close()
