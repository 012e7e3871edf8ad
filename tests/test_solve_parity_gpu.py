"""GPU parity of the wake solve (interface mode = FlorisInterface.update_command, interface.py:557-586) against the
CPU oracle on identical layouts, winds and yaw commands.  FP64: <=1e-9 relative; FP32: <=1e-4 relative (BASELINE.json
north_star); turbine sort order bit-exact."""
import numpy as np
import pytest

from tests._util import CONFIG_LAYOUTS, host_trig, layout, rel_err, sample_winds

pytestmark = pytest.mark.gpu

TOL = {"f64": 1e-9, "f32": 1e-4}


def _run(name, B, precision, kernel, seed):
    import torch

    from oracle import c_oracle
    from wfcrl_b200.backend import FlorisBatch

    lx, ly = layout(name)
    T = len(lx)
    ws, wd = sample_winds(B, seed, tie_every=7)
    rng = np.random.default_rng(seed + 1)
    yaw = rng.uniform(-40, 40, (B, T)).astype(np.float32).astype(np.float64)
    yaw[1] = 0.0
    yaw[2, ::2] = 0.0   # mixed: the kernel's unyawed-source branch next to yawed sources
    yaw[3, 1:] = 0.0    # one yawed turbine (the reference notebook's pattern)
    fb =FlorisBatch(lx, ly, B, precision=precision, kernel=kernel, max_iter=10)
    fb.reset(ws, wd, host_trig=True, warmup_solves=0)
    out = fb.update_command(torch.as_tensor(yaw, device="cuda"))
    torch.cuda.synchronize()
    got = {k: v.cpu().numpy().astype(np.float64) for k, v in out.items()}
    order = fb.get_state("order")
    c, s = host_trig(wd)
    ref = c_oracle.solve_batch(lx, ly, ws, wd, yaw, cs=np.stack([c, s], 1))
    fb.close()
    return got, order, ref, ws, wd


@pytest.mark.parametrize("name", CONFIG_LAYOUTS)
@pytest.mark.parametrize("precision,kernel", [("f64", "basic"), ("f64", "fast"), ("f32", "basic"), ("f32", "fast")])
def test_solve_matches_oracle(cuda_device, name, precision, kernel):
    B = 48
    got, order, ref, ws, wd = _run(name, B, precision, kernel, seed=3)
    tol = TOL[precision]
    assert np.array_equal(order, ref["order"]), "turbine sort order must be bit-exact"
    assert rel_err(got["power"], ref["power_W"], 1.0) <= tol
    assert rel_err(got["wind_speed"], ref["ws_local"], 1e-3) <= tol
    assert rel_err(got["wind_direction"], ref["wd_local"], 1.0) <= tol
    loads_ref = np.stack([ref["ti"], ref["std_u"], ref["std_v"], ref["std_w"]], -1) * 1e7
    # FP32 loads: 2e-4 relative plus an absolute 5e-6 (x1e7): std of v / w are small differences of O(10) m/s values, so
    # their error is absolute (~1e-6 m/s); the basic FP32 kernel (plain transcription, not the product path) keeps 2e-3
    if precision == "f64":
        assert rel_err(got["load"], loads_ref, 1e4) <= tol
    elif kernel == "fast":
        assert np.all(np.abs(got["load"] - loads_ref) <= 2e-4 * np.abs(loads_ref) + 50.0)
    else:
        assert rel_err(got["load"], loads_ref, 1e4) <= 2e-3
    assert np.allclose(got["freewind"][:, 0], ws) and np.allclose(got["freewind"][:, 1], wd)


def test_solve_matches_numpy_oracle_and_golden(cuda_device):
    """Directly against the numpy oracle on the reference's notebook vector (examples/demo.ipynb:137-138)."""
    import json
    import os

    import torch

    from wfcrl_b200.backend import FlorisBatch

    here = os.path.dirname(os.path.abspath(__file__))
    kat = json.load(open(os.path.join(here, "golden", "kat1_ablaincourt.json")))
    lx, ly = layout("Ablaincourt_")
    fb = FlorisBatch(lx, ly, 1, precision="f64", max_iter=10)
    out = fb.reset(kat["wind_speed"], kat["wind_direction"], host_trig=True, warmup_solves=1)
    torch.cuda.synchronize()
    ws_l = out["wind_speed"].cpu().numpy()[0]
    wd_l = out["wind_direction"].cpu().numpy()[0]
    assert np.max(np.abs(ws_l - np.array(kat["local_wind_speed"]))) < 2e-8
    assert np.max(np.abs(wd_l - np.array(kat["local_wind_direction"]))) < 2e-8
    fb.close()


def test_generic_fast_kernel_matches_specialised(cuda_device, monkeypatch):
    """The tuned kernel exists in two instantiations: model constants baked in as immediates (default model) and read
    from kernel parameters / shared memory (any other model).  Both must give the same results (bitwise)."""
    import torch

    from wfcrl_b200.backend import FlorisBatch

    lx, ly = layout("Turb_TCRWP_")
    B, T = 16, len(lx)
    ws, wd = sample_winds(B, 4)
    yaw = torch.as_tensor(np.random.default_rng(0).uniform(-30, 30, (B, T)), device="cuda")
    outs = []
    for forced in (False, True):
        if forced:
            monkeypatch.setenv("WFCRL_B200_NO_BAKED", "1")
        fb = FlorisBatch(lx, ly, B, precision="f32", kernel="fast", max_iter=10)
        fb.reset(ws, wd, warmup_solves=0)
        out = fb.update_command(yaw)
        torch.cuda.synchronize()
        outs.append({k: v.clone() for k, v in out.items()})
        fb.close()
    for k in ("power", "wind_speed", "wind_direction", "load"):
        assert torch.allclose(outs[0][k], outs[1][k], rtol=2e-6, atol=1e-6), k


@pytest.mark.parametrize("precision,kernel", [("f64", "basic"), ("f64", "fast"), ("f32", "fast")])
def test_sampled_turbulence_intensity_per_env(cuda_device, precision, kernel):
    """BASELINE.json configs[2]: wind speed / direction / TI sampled per env (TI ~ U(0.04, 0.12) is an extension of the
    reference, whose case.yaml fixes 0.06): the per-env ambient TI must reach every place FLORIS uses it."""
    import torch

    from oracle import c_oracle
    from wfcrl_b200.backend import FlorisBatch

    lx, ly = layout("Turb16_TCRWP_")
    B, T = 24, len(lx)
    ws, wd = sample_winds(B, 12)
    rng = np.random.default_rng(13)
    ti = rng.uniform(0.04, 0.12, B)
    yaw = rng.uniform(-30, 30, (B, T)).astype(np.float32).astype(np.float64)
    fb = FlorisBatch(lx, ly, B, precision=precision, kernel=kernel, max_iter=10)
    fb.reset(ws, wd, host_trig=True, warmup_solves=0)
    fb.set_turbulence_intensity(torch.as_tensor(ti, device="cuda"))
    out = fb.update_command(torch.as_tensor(yaw, device="cuda"))
    torch.cuda.synchronize()
    p = out["power"].double().cpu().numpy()
    tiout = out["load"].double().cpu().numpy()[..., 0] / 1e7
    c, s = host_trig(wd)
    tol = TOL[precision]
    for b in range(B):
        ref = c_oracle.solve(lx, ly, ws[b], wd[b], yaw[b], cs=(c[b], s[b]), ti_ambient=ti[b])
        assert rel_err(p[b], ref.power_W, 1.0) <= tol, (b, precision)
        assert rel_err(tiout[b], ref.ti, 1e-3) <= max(tol, 1e-9)
    assert np.allclose(fb.get_state("ti_ambient"), ti)
    fb.close()


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_five_by_five_rotor_grid_basic_kernel(cuda_device, precision):
    """SURVEY 8f row 4 (case.yaml:16 `turbine_grid_points`): a 5x5 rotor grid through the basic kernels against the numpy
    oracle (which is generic in the grid size; the C oracle and the tuned kernels are 3x3 only and must refuse)."""
    import torch

    from oracle import floris_oracle
    from wfcrl_b200 import _lib
    from wfcrl_b200.backend import FlorisBatch

    lx, ly = layout("Turb6_Row2_")
    B, T = 6, len(lx)
    ws, wd = sample_winds(B, 31, tie_every=3)
    yaw = np.random.default_rng(32).uniform(-30, 30, (B, T)).astype(np.float32).astype(np.float64)
    fb = FlorisBatch(lx, ly, B, precision=precision, kernel="basic", max_iter=10, config_overrides={"turbine_grid_points": 5})
    fb.reset(ws, wd, host_trig=True, warmup_solves=0)
    out = fb.update_command(torch.as_tensor(yaw, device="cuda"))
    torch.cuda.synchronize()
    got = {k: v.double().cpu().numpy() for k, v in out.items()}
    order = fb.get_state("order")
    fb.close()
    tol = 1e-9 if precision == "f64" else 1e-4
    floris_oracle.CASE["grid_points"] = 5
    try:
        for b in range(B):
            c, s = host_trig(wd[b])
            ref = floris_oracle.solve(lx, ly, ws[b], wd[b], yaw[b], cs=(c, s))
            assert np.array_equal(order[b], ref.order)
            assert rel_err(got["power"][b], ref.power_W, 1.0) <= tol, (b, precision)
            assert rel_err(got["wind_speed"][b], ref.ws_local, 1e-3) <= tol
            assert rel_err(got["wind_direction"][b], ref.wd_local, 1.0) <= tol
            assert rel_err(got["load"][b, :, 0] / 1e7, ref.ti, 1e-3) <= max(tol, 1e-9)
            assert rel_err(got["load"][b, :, 1] / 1e7, ref.std_u, 1e-3) <= (tol if precision == "f64" else 2e-3)
    finally:
        floris_oracle.CASE["grid_points"] = 3
    with pytest.raises(_lib.WfError, match="WF_KERNEL_BASIC"):
        FlorisBatch(lx, ly, B, precision=precision, kernel="fast", config_overrides={"turbine_grid_points": 5})
