"""GPU edge cases: single turbine, odd turbine counts, the maximum supported farm (128 turbines), a batch of one,
exact x-ties, wind from every quadrant, wind speeds at the table's ends, invalid arguments."""
import numpy as np
import pytest

from oracle import c_oracle
from tests._util import host_trig, layout, rel_err

pytestmark = pytest.mark.gpu


def _check(lx, ly, ws, wd, yaw, precision, kernel, tol):
    import torch

    from wfcrl_b200.backend import FlorisBatch

    B = len(ws)
    fb = FlorisBatch(lx, ly, B, precision=precision, kernel=kernel, max_iter=10)
    fb.reset(ws, wd, host_trig=True, warmup_solves=0)
    out = fb.update_command(torch.as_tensor(yaw, device="cuda"))
    torch.cuda.synchronize()
    c, s = host_trig(wd)
    ref = c_oracle.solve_batch(lx, ly, ws, wd, yaw, cs=np.stack([c, s], 1))
    assert np.array_equal(fb.get_state("order"), ref["order"])
    p = out["power"].double().cpu().numpy()
    assert rel_err(p, ref["power_W"], 1.0) <= tol, (precision, kernel)
    assert rel_err(out["wind_speed"].double().cpu().numpy(), ref["ws_local"], 1e-3) <= tol
    assert rel_err(out["load"].double().cpu().numpy()[..., 0], ref["ti"] * 1e7, 1e3) <= max(tol, 1e-9)
    fb.close()


MODES = [("f64", "basic", 1e-9), ("f64", "fast", 1e-9), ("f32", "fast", 1e-4)]


@pytest.mark.parametrize("precision,kernel,tol", MODES)
@pytest.mark.parametrize("name", ["Turb1_Row1_", "Turb2_Row1_", "Turb3_Row1_", "Turb12_Row1_", "HornsRev2_", "WMR_",
                                  "Ormonde_"])
def test_small_and_odd_layouts(cuda_device, name, precision, kernel, tol):
    lx, ly = layout(name)
    T = len(lx)
    rng = np.random.default_rng(T)
    B = 5
    ws = np.array([3.0, 8.0, 11.4, 24.9, 9.3])          # cut-in, rated region, near cut-out
    wd = np.array([270.0, 90.0, 0.0, 181.3, 315.0])     # exact row alignment (ties), reversed, cross, generic
    yaw = rng.uniform(-40, 40, (B, T)).astype(np.float32).astype(np.float64)
    yaw[0] = 0.0
    _check(lx, ly, ws, wd, yaw, precision, kernel, tol)


@pytest.mark.parametrize("precision,kernel,tol", MODES)
def test_maximum_farm_128_turbines(cuda_device, precision, kernel, tol):
    rng = np.random.default_rng(128)
    gx, gy = np.meshgrid(np.arange(16) * 700.0, np.arange(8) * 600.0)
    lx = (gx + rng.uniform(-40, 40, gx.shape)).ravel()
    ly = (gy + rng.uniform(-40, 40, gy.shape)).ravel()
    B = 4
    ws = np.array([7.0, 9.0, 12.0, 15.0])
    wd = np.array([268.0, 275.0, 250.0, 290.0])
    yaw = rng.uniform(-30, 30, (B, 128)).astype(np.float32).astype(np.float64)
    _check(lx, ly, ws, wd, yaw, precision, kernel, tol)


def test_batch_of_one_and_outside_table(cuda_device):
    """B = 1; wind speeds outside the power table's useful range give zero power, not NaN."""
    import torch

    from wfcrl_b200.backend import FlorisBatch

    lx, ly = layout("Turb3_Row1_")
    for precision, kernel in (("f64", "fast"), ("f32", "fast")):
        fb = FlorisBatch(lx, ly, 1, precision=precision, kernel=kernel, max_iter=10)
        out = fb.reset(26.0, 270.0)  # above cut-out (25.02 m/s): Cp = 0
        torch.cuda.synchronize()
        assert torch.isfinite(out["wind_speed"]).all()
        out = fb.step(torch.zeros(1, 3, device="cuda"))
        torch.cuda.synchronize()
        assert float(out["power"].abs().max()) < 1e-6 and torch.isfinite(out["reward"]).all()
        fb.close()


def test_invalid_arguments(cuda_device):
    import torch

    from wfcrl_b200 import _lib
    from wfcrl_b200.backend import FlorisBatch

    lx, ly = layout("Turb3_Row1_")
    with pytest.raises(_lib.WfError):
        FlorisBatch(lx, ly, 0)
    with pytest.raises(_lib.WfError):
        FlorisBatch(np.zeros(129), np.zeros(129), 2)
    with pytest.raises(_lib.WfError):
        FlorisBatch(lx, ly, 2, yaw_bounds=(10.0, -10.0, 1.0))
    with pytest.raises(_lib.WfError):
        FlorisBatch(lx, ly, 2, config_overrides={"wind_veer": 3.0})
    fb = FlorisBatch(lx, ly, 2)
    with pytest.raises(_lib.WfError):
        fb.reset([8.0], [270.0], env_ids=[5])
    with pytest.raises(ValueError):
        fb.step(torch.zeros(2, 4, device="cuda"))
    with pytest.raises(TypeError):
        fb.step(torch.zeros(2, 3, device="cuda", dtype=torch.float64))
    with pytest.raises(KeyError):
        fb.get_state("nope")
    fb.close()
