"""GPU tests of the host-side mirror of the reference API: ``envs.make`` (Gymnasium / PettingZoo-AEC single envs through
the FlorisInterface drop-in) and ``envs.make_vec`` (batched), compared with the env-semantics oracle."""
import numpy as np
import pytest

from oracle import c_oracle, env_oracle
from tests._util import layout

pytestmark = pytest.mark.gpu


def test_make_single_env_reproduces_notebook_flow(cuda_device):
    """examples/demo.ipynb cells 5-16 with the B200 backend: spaces, reset observation, 69-step episode, history."""
    from wfcrl_b200 import environments as envs

    env = envs.make("Ablaincourt_Floris", max_num_steps=70)
    assert env.num_turbines == 7
    assert "yaw" in env.action_space and env.action_space["yaw"].shape == (7,)
    assert float(env.action_space["yaw"].low[0]) == -5.0 and env.action_space["yaw"].dtype == np.float32
    assert list(env.observation_space.keys()) == ["yaw", "freewind_measurements", "wind_speed", "wind_direction"]
    obs = env.reset(options={"wind_speed": 6.48958384, "wind_direction": 266.363907})
    assert np.max(np.abs(obs["wind_speed"] - [6.46819497, 4.58929161, 6.46702757, 6.21243961, 6.20072934, 6.1100638,
                                              5.76785291])) < 2e-8
    assert np.max(np.abs(obs["wind_direction"] - [266.64538262, 267.04575667, 266.77944635, 266.86120544, 266.89071421,
                                                  266.92108378, 266.99680007])) < 2e-8
    lx, ly = layout("Ablaincourt_")
    ref = env_oracle.EnvOracle(lx, ly, solver=c_oracle.solve, max_num_steps=70)
    obs = env.reset(seed=3)
    robs = ref.reset(seed=3)
    assert np.allclose(obs["freewind_measurements"], robs["freewind_measurements"], rtol=0, atol=0)
    total, rtotal, i, done = 0.0, 0.0, 0, False
    while not done:
        a = np.zeros(7)
        if i % 5 == 0:
            a[int(i / 5 % 7)] = -5.0
        obs, reward, term, trunc, info = env.step({"yaw": a.copy()})
        robs, rreward, _t, rtrunc, rinfo = ref.step({"yaw": a.copy()})
        assert np.array_equal(obs["yaw"], robs["yaw"]) and obs["yaw"].dtype == np.float32
        assert abs(reward[0] - rreward[0]) < 1e-9 * abs(rreward[0])
        assert np.allclose(info["power"], rinfo["power"], rtol=1e-9, atol=0)
        assert np.allclose(info["load"], rinfo["load"], rtol=1e-8, atol=1e-12)
        assert trunc == rtrunc and term is False
        total += reward
        rtotal += rreward
        i += 1
        done = term or trunc
    assert i == 69 and len(env.history["observation"]) == 69 and len(env.history["power"]) == 69
    assert abs(total[0] - rtotal[0]) < 1e-8


def test_make_decentralized_env_example_floris(cuda_device):
    """examples/example_floris.py: Dec_Ablaincourt_Floris, StepPercentage, load_coef=1, AEC loop until all agents done."""
    from wfcrl_b200 import environments as envs
    from wfcrl_b200.rewards import StepPercentage

    env = envs.make("Dec_Ablaincourt_Floris", max_num_steps=30, reward_shaper=StepPercentage(), load_coef=1)
    lx, ly = layout("Ablaincourt_")
    ref = env_oracle.MAEnvOracle(lx, ly, solver=c_oracle.solve, max_num_steps=30, load_coef=1,
                                 reward_shaper=env_oracle.StepPercentage())

    def policy(agent, i):
        if agent == "turbine_1" and i == 20:
            return {"yaw": np.array([15.0])}
        if i % 3 == 0:
            return {"yaw": np.array([-4.0])}
        return {"yaw": np.array([0.0])}

    def run(e):
        e.reset(options={"wind_speed": 9.0, "wind_direction": 268.0})
        r = {a: 0 for a in e.possible_agents}
        done = {a: False for a in e.possible_agents}
        n = {a: 0 for a in e.possible_agents}
        for agent in e.agent_iter():
            obs, reward, term, trunc, info = e.last()
            done[agent] = done[agent] or term or trunc
            r[agent] += reward
            if done[agent]:
                action = None
            else:
                action = policy(agent, n[agent])
                n[agent] += 1
            e.step(action)
        return r, n

    r, n = run(env)
    rr, rn = run(ref)
    assert n == rn and all(v == 29 for v in n.values())
    for a in r:
        assert abs(float(np.asarray(r[a]).reshape(-1)[0]) - float(np.asarray(rr[a]).reshape(-1)[0])) < 1e-7
    assert set(env.history["turbine_1"].keys()) == {"observation", "reward", "load", "power"}
    assert len(env.history["turbine_3"]["power"]) > 0


def test_control_validation_errors(cuda_device):
    from wfcrl_b200 import environments as envs

    with pytest.raises(ValueError):
        envs.make("Ablaincourt_Floris", controls={"pitch": (0, 45, 1)})
    with pytest.raises(ValueError):
        envs.make("Ablaincourt_Floris", controls={"yaw": (20, -20, 1)})
    with pytest.raises(ValueError):
        envs.make("NoSuchFarm_Floris")
    with pytest.warns(UserWarning):
        env = envs.make("Turb3_Row1_Floris", controls={"yaw": (-20, 20)}, log=False)
    assert env.controls["yaw"] == (-20, 20, 1)


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_make_vec_matches_oracle_with_autoreset(cuda_device, precision):
    import torch

    from wfcrl_b200 import environments as envs

    B, steps, max_steps = 6, 9, 6
    env = envs.make_vec("Turb6_Row2_Floris", B, precision=precision, max_num_steps=max_steps, exact_host_trig=True)
    obs = env.reset(seed=100)
    lx, ly = layout("Turb6_Row2_")
    refs = [env_oracle.EnvOracle(lx, ly, solver=c_oracle.solve, max_num_steps=max_steps) for _ in range(B)]
    robs = [r.reset(seed=100 + b) for b, r in enumerate(refs)]
    tol = 1e-9 if precision == "f64" else 1e-4
    fw = obs["freewind_measurements"].double().cpu().numpy()
    for b in range(B):
        assert np.allclose(fw[b], robs[b]["freewind_measurements"], rtol=1e-7 if precision == "f32" else 0, atol=0)
    rng = np.random.default_rng(0)
    n_trunc = 0
    first_episode_return = np.zeros(B)
    for k in range(steps):
        a = rng.uniform(-5, 5, (B, 6)).astype(np.float32)
        obs, reward, term, trunc, info = env.step(torch.as_tensor(a, device="cuda"))
        rew = reward.double().cpu().numpy()
        tr = trunc.cpu().numpy()
        assert not term.any()
        if n_trunc == 0:
            first_episode_return += rew
        if k < max_steps - 1:
            for b, r in enumerate(refs):
                o, rr, _t, rtr, _i = r.step({"yaw": a[b].copy()})
                assert abs(rew[b] - rr[0]) <= tol * max(1, abs(rr[0]))
                assert bool(tr[b]) == bool(rtr)
                if not rtr:
                    assert np.array_equal(obs["yaw"][b].cpu().numpy().astype(np.float32), o["yaw"])
        if tr.any():
            n_trunc += 1
            assert tr.all() and "final_observation" in info
            # auto-reset: yaw back to zero, a fresh episode has started
            assert float(obs["yaw"].abs().max()) == 0.0
    assert n_trunc == 1
    stats = env.episode_statistics()   # from the sums the step kernels keep per env (no per-step host bookkeeping)
    assert stats["episodes"] == B and stats["length_mean"] == max_steps - 1
    assert abs(stats["return_mean"] - first_episode_return.mean()) <= 1e-12 * abs(first_episode_return.mean())
    assert np.allclose(env.episode_lengths.cpu().numpy(), steps - (max_steps - 1))   # the running (second) episode
    env.close()


def test_make_vec_decentralized_and_time_series(cuda_device):
    import torch

    from wfcrl_b200 import environments as envs

    B = 3
    env = envs.make_vec("Dec_Ablaincourt_Floris", B, precision="f64", max_num_steps=10)
    obs = env.reset(options={"wind_speed": 8.0, "wind_direction": 270.0})
    assert set(obs.keys()) == set(env.possible_agents) and set(obs["turbine_1"].keys()) == {"yaw", "wind_speed", "wind_direction"}
    acts = {a: {"yaw": torch.full((B,), 5.0)} for a in env.possible_agents}
    for _ in range(3):
        obs, rew, term, trunc, info = env.step(acts)
    # stale accumulator for non-last agents: 5, 10, 15 ; last agent: 5, 5, 10 (see test_env_oracle)
    assert float(obs["turbine_1"]["yaw"][0]) == 15.0 and float(obs["turbine_7"]["yaw"][0]) == 10.0
    assert rew["turbine_1"].shape == (B,) and info["turbine_2"]["load"].shape == (B, 4)
    env.close()
    # wind time series: wind changes before every solve; reward normalised with the PREVIOUS state's speed
    series = np.array([[8.0, 270.0], [9.0, 265.0], [7.0, 275.0]])
    env = envs.make_vec("Turb3_Row1_Floris", 2, precision="f64", max_num_steps=10, wind_time_series=series)
    np.random.seed(0)
    env.reset()
    lx, ly = layout("Turb3_Row1_")
    pos = env._series_pos.cpu().numpy().copy()
    prev_ws = series[pos, 0]
    for k in range(4):
        obs, reward, term, trunc, info = env.step(torch.zeros(2, 3))
        pos = (pos + 1) % 3
        for b in range(2):
            sol = c_oracle.solve(lx, ly, series[pos[b], 0], series[pos[b], 1], np.zeros(3))
            expect = np.mean(sol.power_W / 1e6 * 1e3 / prev_ws[b] ** 3) - 0.1 * np.mean(
                np.abs(np.stack([sol.ti, sol.std_u, sol.std_v, sol.std_w], 1)))
            assert abs(float(reward[b]) - expect) < 1e-9 * abs(expect)
            assert abs(float(obs["freewind_measurements"][b, 0]) - series[pos[b], 0]) < 1e-12
        prev_ws = series[pos, 0]
    env.close()


def test_sharded_batches_equal_single_batch(cuda_device):
    """env ids are global, so a batch split into shards gives identical per-env trajectories."""
    import torch

    from wfcrl_b200 import environments as envs
    from wfcrl_b200.dist import shard_range

    B, T = 10, 16
    full = envs.make_vec("Turb16_Row5_Floris", B, precision="f32", max_num_steps=20)
    full.reset(seed=7)
    shards = []
    for r in range(3):
        lo, hi = shard_range(B, r, 3)
        e = envs.make_vec("Turb16_Row5_Floris", hi - lo, precision="f32", max_num_steps=20, env_id_offset=lo)
        e.reset(seed=7)
        shards.append((lo, hi, e))
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1)
    for _ in range(5):
        a = torch.rand(B, T, device="cuda", generator=gen) * 10 - 5
        obs, rew, *_ = full.step(a)
        for lo, hi, e in shards:
            o2, r2, *_ = e.step(a[lo:hi].contiguous())
            assert torch.equal(r2, rew[lo:hi]) and torch.equal(o2["wind_speed"], obs["wind_speed"][lo:hi])
    full.close()
    for *_x, e in shards:
        e.close()


def test_make_vec_sampled_turbulence_intensity(cuda_device):
    """BASELINE configs[2] flavour: TI sampled per env at reset; start observation matches the oracle run with that TI."""
    from wfcrl_b200 import environments as envs

    B = 5
    env = envs.make_vec("Turb16_TCRWP_Floris", B, precision="f64", max_num_steps=10, turbulence_intensity_range=(0.04, 0.12))
    obs = env.reset(seed=50)
    ti = env.backend.get_state("ti_ambient")
    assert np.all((ti >= 0.04) & (ti <= 0.12)) and len(np.unique(ti)) == B
    lx, ly = layout("Turb16_TCRWP_")
    fw = obs["freewind_measurements"].cpu().numpy()
    for b in range(B):
        ref = c_oracle.solve(lx, ly, fw[b, 0], fw[b, 1], np.zeros(16), ti_ambient=ti[b])
        assert np.allclose(obs["wind_speed"][b].cpu().numpy(), np.clip(ref.ws_local, 3, 28), rtol=1e-9)
    env.close()


def test_single_env_wind_time_series_matches_oracle(cuda_device):
    """Wind time-series mode through the drop-in path (interface.py:503-524, 563): random start offset from numpy's global
    RNG, wind updated before every solve, requested reset winds ignored with a warning."""
    from wfcrl_b200 import environments as envs

    series = np.array([[8.0, 270.0], [9.5, 262.0], [7.2, 281.0], [11.0, 255.0], [6.4, 275.5]])
    lx, ly = layout("Turb3_Row1_")
    np.random.seed(11)
    env = envs.make("Turb3_Row1_Floris", max_num_steps=12, wind_time_series=series, log=False)
    np.random.seed(11)
    ref = env_oracle.EnvOracle(lx, ly, solver=c_oracle.solve, max_num_steps=12, wind_time_series=series)
    np.random.seed(12)  # the start offset of every reset comes from numpy's GLOBAL generator (interface.py:517)
    obs = env.reset(seed=0)
    np.random.seed(12)
    robs = ref.reset(seed=0)
    assert np.allclose(obs["freewind_measurements"], robs["freewind_measurements"])
    rng = np.random.default_rng(1)
    for _ in range(7):  # runs past the end of the 5-row series? no: 5 rows, start offset + 7 draws would exhaust it
        a = rng.uniform(-5, 5, 3).astype(np.float32)
        try:
            r_out = ref.step({"yaw": a.copy()})
        except (StopIteration, RuntimeError):
            with pytest.raises((StopIteration, RuntimeError)):
                env.step({"yaw": a.copy()})
            break
        o, r, _t, tr, info = env.step({"yaw": a.copy()})
        ro, rr, _rt, rtr, rinfo = r_out
        assert np.allclose(o["freewind_measurements"], ro["freewind_measurements"])
        assert abs(r[0] - rr[0]) <= 1e-9 * abs(rr[0]) and tr == rtr
        assert np.allclose(info["power"], rinfo["power"], rtol=1e-9)


def test_reset_sampled_draws_match_checker_and_ignore_sharding(cuda_device):
    """wf_reset_sampled: the library's counter-based reset sampler reproduces the checker value by
    value, advances the per-env episode counter only for the selected envs, and depends on global env ids only."""
    import torch

    from wfcrl_b200.backend import FlorisBatch

    lx, ly = layout("Turb6_Row2_")
    B, seed = 12, 0x1234_5678_9ABC
    fb = FlorisBatch(lx, ly, B, precision="f64", kernel="fast", max_iter=10)
    out = fb.reset_sampled(None, seed, env_id_offset=0)
    fw = out["freewind"].cpu().numpy().copy()
    want = np.array([env_oracle.sampled_reset_wind(seed, g, 0) for g in range(B)])
    assert np.allclose(fw, want, rtol=1e-13, atol=0)
    assert np.array_equal(fb.get_state("ws"), fw[:, 0]) and list(fb.get_state("episode")) == [1] * B
    sol = c_oracle.solve(lx, ly, fw[3, 0], fw[3, 1], np.zeros(6))
    assert np.allclose(out["wind_speed"][3].cpu().numpy(), sol.ws_local, rtol=1e-9)   # warm-up solve ran on the new wind
    assert list(fb.get_state("num_iter")) == [1] * B
    # masked second episode with TI: only the selected envs move on
    mask = torch.zeros(B, dtype=torch.uint8, device="cuda")
    mask[[2, 5]] = 1
    out = fb.reset_sampled(mask, seed, 0, turbulence_intensity_range=(0.04, 0.12))
    fw2 = out["freewind"].cpu().numpy()
    ti = fb.get_state("ti_ambient")
    for g in range(B):
        if g in (2, 5):
            ws, wd, t = env_oracle.sampled_reset_wind(seed, g, 1, ti_range=(0.04, 0.12))
            assert np.allclose(fw2[g], [ws, wd], rtol=1e-13) and abs(ti[g] - t) < 1e-15
        else:
            assert np.array_equal(fw2[g], fw[g]) and ti[g] == 0.06
    assert list(fb.get_state("episode")) == [2 if g in (2, 5) else 1 for g in range(B)]
    fb.close()
    # two shards with global offsets draw the same bits as the single batch
    for lo, hi in ((0, 5), (5, 12)):
        sh = FlorisBatch(lx, ly, hi - lo, precision="f64", kernel="fast", max_iter=10)
        o = sh.reset_sampled(None, seed, env_id_offset=lo)
        assert np.array_equal(o["freewind"].cpu().numpy(), fw[lo:hi])
        sh.close()
    # argument validation through the C-ABI
    fb = FlorisBatch(lx, ly, 2, precision="f32", kernel="fast", max_iter=10)
    with pytest.raises(Exception, match="env_id_offset"):
        fb.reset_sampled(None, 1, env_id_offset=-1)
    fb.close()


def test_vec_env_autoreset_trajectories_ignore_sharding(cuda_device):
    """Two episodes with in-loop auto-resets: a 2-shard run reproduces the single-batch run bit for bit (wind of the
    second episode included), for seeded and for unseeded-but-same-key starts."""
    import torch

    from wfcrl_b200 import environments as envs

    B, T, max_steps = 8, 6, 4
    gen = torch.Generator(device="cuda").manual_seed(2)
    acts = [torch.rand(B, T, device="cuda", generator=gen) * 10 - 5 for _ in range(7)]

    def run(envs_and_ranges, seed):
        rows = []
        for e, _lo, _hi in envs_and_ranges:
            if seed is None:
                e._seed = 99
            e.reset(seed=seed)
        for a in acts:
            parts = [e.step(a[lo:hi].contiguous()) for e, lo, hi in envs_and_ranges]
            rows.append((torch.cat([p[0]["freewind_measurements"] for p in parts]).clone(),
                         torch.cat([p[1] for p in parts]).clone(), torch.cat([p[3] for p in parts]).clone()))
        return rows

    for seed in (21, None):
        full = [(envs.make_vec("Turb6_Row2_Floris", B, precision="f32", max_num_steps=max_steps), 0, B)]
        halves = [(envs.make_vec("Turb6_Row2_Floris", 4, precision="f32", max_num_steps=max_steps, env_id_offset=lo), lo, lo + 4)
                  for lo in (0, 4)]
        ra, rb = run(full, seed), run(halves, seed)
        n_trunc = 0
        for (fa, wa, ta), (fb_, wb, tb) in zip(ra, rb):
            assert torch.equal(fa, fb_) and torch.equal(wa, wb) and torch.equal(ta, tb)
            n_trunc += int(ta.all())
        assert n_trunc == 2
        assert not torch.equal(ra[0][0], ra[-1][0])   # the wind did change across episodes
        for e, *_ in full + halves:
            e.close()


def test_vec_log_wrapper_ring_buffers(cuda_device):
    """Device-side history: chronological, wraps around, logs the finished episode's last rows on
    auto-reset steps; works for the centralised and the decentralised batch."""
    import torch

    from wfcrl_b200 import environments as envs

    B, T = 4, 6
    env = envs.make_vec("Turb6_Row2_Floris", B, precision="f64", max_num_steps=4, log=5)
    plain = envs.make_vec("Turb6_Row2_Floris", B, precision="f64", max_num_steps=4)
    env.reset(seed=9)
    plain.reset(seed=9)
    gen = torch.Generator(device="cuda").manual_seed(4)
    seen = []
    for k in range(7):
        a = torch.rand(B, T, device="cuda", generator=gen) * 10 - 5
        env.step(a)
        obs, rew, _t, trunc, info = plain.step(a)
        last = info.get("final_observation", obs)
        seen.append((last["yaw"].clone(), rew.clone(), info.get("final_info", info)["power"].clone(), trunc.clone()))
    h = env.history
    assert h["reward"].shape == (5, B) and h["observation"]["yaw"].shape == (5, B, T) and h["load"].shape == (5, B, T, 4)
    for row, (yaw, rew, power, trunc) in enumerate(seen[-5:]):
        assert torch.equal(h["observation"]["yaw"][row], yaw) and torch.equal(h["reward"][row], rew)
        assert torch.equal(h["power"][row], power) and torch.equal(h["truncated"][row], trunc)
    assert h["truncated"][:, 0].tolist() == [True, False, False, True, False]   # steps 2..6, truncation at 2 and 5
    assert float(h["observation"]["yaw"][0].abs().max()) > 0   # the truncation row is the episode's last state, not the restart
    env.close()
    plain.close()
    dec = envs.make_vec("Dec_Turb3_Row1_Floris", 2, precision="f32", max_num_steps=10, log=8)
    dec.reset(seed=1)
    for _ in range(3):
        dec.step(torch.ones(2, 3, device="cuda"))
    assert dec.history["reward"].shape == (3, 2) and dec.history["observation"]["yaw"][-1, 0].tolist() == [3.0, 3.0, 3.0]
    dec.close()
