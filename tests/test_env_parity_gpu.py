"""GPU parity of the fused env step (constraint + yaw transition + solve + measures + reward + truncation) against the
env-semantics oracle, step by step over whole episodes.  Integer/boolean semantics (yaw trajectory in float32, zeroed
actions, truncation step) must be bit-exact; power/reward <=1e-9 (FP64) or <=1e-4 (FP32) relative."""
import numpy as np
import pytest

from oracle import c_oracle, env_oracle
from tests._util import layout, sample_winds

pytestmark = pytest.mark.gpu

TOL = {"f64": 1e-9, "f32": 1e-4}


def _oracle_envs(name, B, ws, wd, **kw):
    lx, ly = layout(name)
    envs = [env_oracle.EnvOracle(lx, ly, solver=c_oracle.solve, **kw) for _ in range(B)]
    obs = [e.reset(options={"wind_speed": ws[b], "wind_direction": wd[b]}) for b, e in enumerate(envs)]
    return envs, obs


@pytest.mark.parametrize("name,precision,kernel,steps", [
    ("Turb6_Row2_", "f64", "basic", 30), ("Turb6_Row2_", "f64", "fast", 30), ("Turb6_Row2_", "f32", "fast", 30),
    ("Ablaincourt_", "f64", "fast", 25), ("Turb32_Row5_", "f64", "fast", 12), ("HornsRev1_", "f64", "fast", 6),
    ("Ablaincourt_", "f64", "basic", 25), ("Ablaincourt_", "f32", "fast", 25), ("Ablaincourt_", "f32", "basic", 10),
    ("Turb_TCRWP_", "f32", "fast", 12), ("Turb32_Row5_", "f64", "basic", 12), ("HornsRev1_", "f64", "basic", 6),
    ("HornsRev1_", "f32", "fast", 6),
])
def test_env_step_matches_oracle(cuda_device, name, precision, kernel, steps):
    import torch

    from wfcrl_b200.backend import FlorisBatch

    lx, ly = layout(name)
    T, B = len(lx), 12
    max_num_steps = steps - 2  # exercise truncation inside the run (truncates at step max_num_steps-1)
    ws, wd = sample_winds(B, seed=5, tie_every=4)
    ws[1] = 2.0  # below the observation-space bound: start state is clipped, reward normalisation uses the clipped value
    fb = FlorisBatch(lx, ly, B, precision=precision, kernel=kernel, max_iter=max_num_steps, load_coef=0.1)
    start = fb.reset(ws, wd, host_trig=True)
    torch.cuda.synchronize()
    envs, obs0 = _oracle_envs(name, B, ws, wd, max_num_steps=max_num_steps, load_coef=0.1)
    tol = TOL[precision]
    g = {k: v.double().cpu().numpy() for k, v in start.items()}
    for b in range(B):
        assert np.allclose(g["wind_speed"][b], obs0[b]["wind_speed"], rtol=tol, atol=0)
        assert np.allclose(g["wind_direction"][b], obs0[b]["wind_direction"], rtol=tol, atol=0)
        assert np.allclose(g["freewind"][b], obs0[b]["freewind_measurements"], rtol=1e-7)
    rng = np.random.default_rng(9)
    done = np.zeros(B, dtype=bool)
    for k in range(steps):
        a = rng.uniform(-7, 7, (B, T)).astype(np.float32)  # beyond +-5 to exercise the action clip
        a[:, 0] = 5.0  # saturate one turbine so that the actuation constraint and the +-40 clip trigger
        out = fb.step(torch.as_tensor(a, device="cuda"))
        torch.cuda.synchronize()
        g = {k2: v.double().cpu().numpy() for k2, v in out.items()}
        for b, e in enumerate(envs):
            if done[b]:
                continue
            obs, r, term, trunc, info = e.step({"yaw": a[b].copy()})
            assert np.array_equal(g["yaw"][b].astype(np.float32), obs["yaw"]), (k, b)  # bit-exact float32 yaw state
            assert bool(g["truncated"][b]) == bool(trunc) and not term, (k, b)
            perr = np.max(np.abs(g["power"][b] - info["power"]) / np.maximum(info["power"], 1e-3))
            assert perr <= tol, (k, b, perr)
            assert abs(g["reward"][b] - r[0]) <= tol * max(1.0, abs(r[0])), (k, b)
            assert np.allclose(g["wind_speed"][b], obs["wind_speed"], rtol=tol, atol=0), (k, b)
            assert np.allclose(g["wind_direction"][b], obs["wind_direction"], rtol=tol, atol=0), (k, b)
            ltol = tol if precision == "f64" else 2e-3
            assert np.max(np.abs(g["load"][b] - info["load"]) / np.maximum(np.abs(info["load"]), 1e-3)) <= ltol, (k, b)
            done[b] |= bool(trunc)
        if done.all():
            assert k == max_num_steps - 2  # reset consumed one iteration: truncation on step max_num_steps-1 (0-based k)
            break
    assert done.all()
    acc = fb.get_state("acc")
    for b, e in enumerate(envs):
        assert np.array_equal(acc[b], e.mdp._acc["yaw"])
    fb.close()


@pytest.mark.parametrize("precision,kernel", [("f64", "basic"), ("f64", "fast"), ("f32", "fast")])
def test_multi_agent_constraint_discrete_and_shapers(cuda_device, precision, kernel):
    import torch

    from wfcrl_b200.backend import FlorisBatch

    lx, ly = layout("Ablaincourt_")
    T, B, steps = len(lx), 4, 14
    ws, wd = sample_winds(B, seed=2)
    tol = TOL[precision]
    # (a) multi-agent staleness of the actuation constraint + StepPercentage shaper (examples/example_floris.py)
    fb = FlorisBatch(lx, ly, B, precision=precision, kernel=kernel, max_iter=50, load_coef=1.0, multi_agent=True,
                     reward_shaper="step")
    fb.reset(ws, wd)
    envs = [env_oracle.MAEnvOracle(lx, ly, solver=c_oracle.solve, max_num_steps=50, load_coef=1.0,
                                   reward_shaper=env_oracle.StepPercentage()) for _ in range(B)]
    for b, e in enumerate(envs):
        e.reset(options={"wind_speed": ws[b], "wind_direction": wd[b]})
    rng = np.random.default_rng(3)
    for k in range(steps):
        a = rng.choice([-5.0, 0.0, 5.0], size=(B, T)).astype(np.float32)
        out = fb.step(torch.as_tensor(a, device="cuda"))
        torch.cuda.synchronize()
        yaw = out["yaw"].cpu().numpy().astype(np.float32)
        rew = out["reward"].double().cpu().numpy()
        for b, e in enumerate(envs):
            for j, agent in enumerate(list(e.agents)):
                e.step({"yaw": np.array([a[b, j]])})
            assert np.array_equal(yaw[b], e._state["yaw"]), (k, b)
            r = e.rewards[e.possible_agents[0]][0]
            assert abs(rew[b] - r) <= 50 * tol * max(1.0, abs(r)), (k, b, rew[b], r)  # shaped reward is a small difference
    fb.close()
    # (b) discrete control {0,1,2} -> (a-1)*step (mdp.py:306-310) + ReferencePercentage shaper
    fb = FlorisBatch(lx, ly, B, precision=precision, kernel=kernel, max_iter=50, continuous_control=False,
                     reward_shaper="reference", shaper_reference=2.0, yaw_bounds=(-20.0, 20.0, 3.0))
    fb.reset(ws, wd)
    envs = [env_oracle.EnvOracle(lx, ly, solver=c_oracle.solve, max_num_steps=50, continuous_control=False,
                                 controls={"yaw": (-20, 20, 3)}, reward_shaper=env_oracle.ReferencePercentage(2.0))
            for _ in range(B)]
    for b, e in enumerate(envs):
        e.reset(options={"wind_speed": ws[b], "wind_direction": wd[b]})
    for k in range(8):
        a = rng.integers(0, 3, size=(B, T)).astype(np.float32)
        out = fb.step(torch.as_tensor(a, device="cuda"))
        torch.cuda.synchronize()
        yaw = out["yaw"].cpu().numpy().astype(np.float32)
        rew = out["reward"].double().cpu().numpy()
        for b, e in enumerate(envs):
            obs, r, *_ = e.step({"yaw": a[b].copy()})
            assert np.array_equal(yaw[b], obs["yaw"])
            assert abs(rew[b] - r[0]) <= 5 * tol * max(1.0, abs(r[0]))
    fb.close()


def test_host_path_and_masked_reset(cuda_device):
    """wf_step_host (HOST buffers in/out) gives the same result as the device path; wf_reset_masked restarts episodes."""
    import torch

    from wfcrl_b200.backend import FlorisBatch

    lx, ly = layout("Turb6_Row2_")
    B, T = 8, len(lx)
    ws, wd = sample_winds(B, seed=1)
    fa = FlorisBatch(lx, ly, B, precision="f32", kernel="fast", max_iter=6)
    fbb = FlorisBatch(lx, ly, B, precision="f32", kernel="fast", max_iter=6)   # pinned buffers: zero-copy launch
    fcc = FlorisBatch(lx, ly, B, precision="f32", kernel="fast", max_iter=6)   # pageable action: staged copies
    for f in (fa, fbb, fcc):
        f.reset(ws, wd, host_trig=False)
    rng = np.random.default_rng(0)
    for k in range(5):
        a = rng.uniform(-5, 5, (B, T)).astype(np.float32)
        dev = fa.step(torch.as_tensor(a, device="cuda"))
        host = fbb.step_host(torch.as_tensor(a).pin_memory())
        staged = fcc.step_host(torch.as_tensor(a))
        torch.cuda.synchronize()
        for key in ("yaw", "power", "reward", "truncated", "load", "wind_speed", "wind_direction", "freewind"):
            assert torch.equal(dev[key].cpu(), host[key]), key
            assert torch.equal(dev[key].cpu(), staged[key]), key
    d2h = B * T * 4 * 8 + B * (4 + 8 + 1)
    assert fbb.last_h2d_bytes == B * T * 4 and fbb.last_d2h_bytes == d2h and fcc.last_d2h_bytes == d2h
    fcc.close()
    assert bool(dev["truncated"].all())  # max_iter=6: reset consumed 1, 5 steps -> truncated
    mask = torch.zeros(B, dtype=torch.uint8, device="cuda")
    mask[::2] = 1
    fa.reset_masked(mask, torch.as_tensor(ws, device="cuda"), torch.as_tensor(wd, device="cuda"))
    torch.cuda.synchronize()
    it = fa.get_state("num_iter")
    assert list(it[::2]) == [1] * (B // 2) and list(it[1::2]) == [6] * (B // 2)
    assert np.all(fa.get_state("yaw")[::2] == 0) and np.any(fa.get_state("yaw")[1::2] != 0)
    fa.close()
    fbb.close()


def test_host_path_large_batch_zero_copy_equals_staged(cuda_device, monkeypatch):
    """4096 envs (the staged pipeline splits into chunks): pinned zero-copy == forced staged == device path."""
    import torch

    from wfcrl_b200.backend import FlorisBatch

    lx, ly = layout("Ablaincourt_")
    B, T = 4096 + 37, len(lx)
    ws, wd = sample_winds(B, seed=2)
    fbs = [FlorisBatch(lx, ly, B, precision="f32", kernel="fast", max_iter=100) for _ in range(3)]
    for f in fbs:
        f.reset(ws, wd, host_trig=False)
    gen = torch.Generator().manual_seed(0)
    for k in range(3):
        a = (torch.rand(B, T, generator=gen) * 10 - 5).pin_memory()
        dev = fbs[0].step(a.cuda())
        monkeypatch.setenv("WFCRL_B200_HOST_PATH", "zero_copy")
        zero = fbs[1].step_host(a)
        monkeypatch.setenv("WFCRL_B200_HOST_PATH", "staged")
        staged = fbs[2].step_host(a)
        monkeypatch.delenv("WFCRL_B200_HOST_PATH")
        torch.cuda.synchronize()
        for key in ("yaw", "power", "reward", "truncated", "load", "wind_speed", "wind_direction", "freewind"):
            assert torch.equal(dev[key].cpu(), zero[key]) and torch.equal(zero[key], staged[key]), key
    # per step: one step launch + one FP64 re-solve launch on the zero-copy route; six chunk launches + ONE re-solve launch
    # (shared fix-up list) on the staged one
    assert fbs[1].launch_count() + 5 * 3 == fbs[2].launch_count()
    for f in fbs:
        f.close()


def test_device_trig_geometry_matches_host_within_ulp(cuda_device):
    """Throughput mode computes cosd/sind on the device: rotated coordinates agree with the host ones to ~1 ulp and the
    oracle fed with the device's cos/sin reproduces the FP64 kernel."""
    import torch

    from wfcrl_b200.backend import FlorisBatch

    lx, ly = layout("HornsRev1_")
    B, T = 16, len(lx)
    ws, wd = sample_winds(B, seed=8)
    fb = FlorisBatch(lx, ly, B, precision="f64", max_iter=10)
    fb.reset(ws, wd, host_trig=False, warmup_solves=0)
    cs = fb.get_state("cs")
    dev = ((wd - 270.0) % 360.0 + 360.0) % 360.0
    assert np.max(np.abs(cs[:, 0] - np.cos(np.radians(dev)))) < 5e-16
    assert np.max(np.abs(cs[:, 1] - np.sin(np.radians(dev)))) < 5e-16
    yaw = np.random.default_rng(1).uniform(-30, 30, (B, T)).astype(np.float32).astype(np.float64)
    out = fb.update_command(torch.as_tensor(yaw, device="cuda"))
    torch.cuda.synchronize()
    ref = c_oracle.solve_batch(lx, ly, ws, wd, yaw, cs=cs)
    assert np.array_equal(fb.get_state("order"), ref["order"])
    p = out["power"].cpu().numpy()
    assert np.max(np.abs(p - ref["power_W"]) / np.maximum(ref["power_W"], 1.0)) < 1e-9
    fb.close()


@pytest.mark.parametrize("precision,kernel", [("f64", "fast"), ("f32", "fast")])
def test_checkpoint_resume_is_exact(cuda_device, precision, kernel):
    """state_dict -> new handle -> load_state_dict continues the episode bit for bit."""
    import torch

    from wfcrl_b200.backend import FlorisBatch

    lx, ly = layout("Turb16_Row5_")
    B, T = 32, len(lx)
    ws, wd = sample_winds(B, seed=40)
    rng = np.random.default_rng(41)
    acts = [torch.as_tensor(rng.uniform(-5, 5, (B, T)).astype(np.float32), device="cuda") for _ in range(8)]
    fa = FlorisBatch(lx, ly, B, precision=precision, kernel=kernel, max_iter=7, reward_shaper="step")
    fa.reset(ws, wd, host_trig=False)
    for k in range(4):
        fa.step(acts[k])
    sd = fa.state_dict()
    fbb = FlorisBatch(lx, ly, B, precision=precision, kernel=kernel, max_iter=7, reward_shaper="step")
    fbb.load_state_dict(sd)
    for k in range(4, 8):
        oa = fa.step(acts[k])
        ob = fbb.step(acts[k])
        torch.cuda.synchronize()
        for key in ("yaw", "power", "reward", "truncated", "wind_speed", "wind_direction", "load", "freewind"):
            assert torch.equal(oa[key], ob[key]), (k, key)
    assert np.array_equal(fa.get_state("num_iter"), fbb.get_state("num_iter"))
    assert np.all(fa.get_state("nonfinite") == 0)
    fa.close()
    fbb.close()


def test_nonfinite_guard_counter(cuda_device):
    """A zero wind speed makes the reference arithmetic divide by zero; the guard counter records it per env."""
    import torch

    from wfcrl_b200.backend import FlorisBatch

    lx, ly = layout("Turb3_Row1_")
    fb = FlorisBatch(lx, ly, 2, precision="f32", kernel="fast", max_iter=10)
    fb.reset([8.0, 8.0], [270.0, 270.0])
    fb.set_state("ws_norm", [8.0, 0.0])
    out = fb.step(torch.zeros(2, 3, device="cuda"))
    torch.cuda.synchronize()
    assert list(fb.get_state("nonfinite")) == [0, 1] and not bool(torch.isfinite(out["reward"][1]))
    fb.close()


def test_step_is_cuda_graph_capturable(cuda_device):
    """wf_step is a pure stream operation (no host sync, no allocation): a captured graph of several env steps replays
    to the same bits as eager launches -- the way to drive small layouts, where launch overhead is the step time."""
    import torch

    from wfcrl_b200.backend import FlorisBatch

    lx, ly = layout("Ablaincourt_")
    B, T, K = 64, len(lx), 4
    ws, wd = sample_winds(B, seed=4)
    eager = FlorisBatch(lx, ly, B, precision="f32", kernel="fast", max_iter=1000)
    graphed = FlorisBatch(lx, ly, B, precision="f32", kernel="fast", max_iter=1000)
    for f in (eager, graphed):
        f.reset(ws, wd, host_trig=False)
    gen = torch.Generator(device="cuda").manual_seed(0)
    actions = [torch.rand(B, T, device="cuda", generator=gen) * 10 - 5 for _ in range(K)]
    static = [a.clone() for a in actions]
    rewards = torch.zeros(K, B, device="cuda")
    graphed.step(static[0])          # warm-up outside capture (lazy function attributes)
    eager.step(actions[0])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for k in range(K):
            out = graphed.step(static[k])
            rewards[k].copy_(out["reward"])
    for rep in range(2):             # two replays = 2 K further env steps
        g.replay()
        want = []
        for k in range(K):
            want.append(eager.step(actions[k])["reward"].clone())
        torch.cuda.synchronize()
        assert torch.equal(rewards, torch.stack(want)), rep
    assert torch.equal(graphed.out["yaw"], eager.out["yaw"])
    assert np.array_equal(graphed.get_state("num_iter"), eager.get_state("num_iter"))
    eager.close()
    graphed.close()


def test_host_path_is_ordered_after_async_calls(cuda_device, monkeypatch):
    """wf_step_host runs on the handle's own streams: it must wait for whatever the caller queued asynchronously on ITS stream
    (device-side reset, wind update, device-side step) without an explicit synchronisation in between."""
    import torch

    from wfcrl_b200.backend import FlorisBatch

    lx, ly = layout("Turb32_Row5_")
    B, T = 4096, len(lx)
    gen = torch.Generator().manual_seed(3)
    acts = [(torch.rand(B, T, generator=gen) * 10 - 5).pin_memory() for _ in range(3)]
    results = {}
    for mode in ("racy", "synced"):
        for path in ("staged", "zero_copy"):
            monkeypatch.setenv("WFCRL_B200_HOST_PATH", path)
            fb = FlorisBatch(lx, ly, B, precision="f32", kernel="fast", max_iter=100)
            outs = []
            for k, a in enumerate(acts):
                fb.reset_sampled(None, seed=5 + k, env_id_offset=0)           # asynchronous: sampler, state, geometry, warm-up
                fb.step(a.cuda(non_blocking=True))                             # asynchronous device-side step
                if mode == "synced":
                    torch.cuda.synchronize()
                outs.append({key: v.clone() for key, v in fb.step_host(a).items()})
            results[(mode, path)] = outs
            fb.close()
    monkeypatch.delenv("WFCRL_B200_HOST_PATH")
    for path in ("staged", "zero_copy"):
        for a, b in zip(results[("racy", path)], results[("synced", path)]):
            for key in a:
                assert torch.equal(a[key], b[key]), (path, key)
