"""CPU tests: the oracle against the reference's golden vectors."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle, floris_oracle
from tests._util import CONFIG_LAYOUTS, host_trig, layout, sample_winds

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden(name):
    with open(os.path.join(HERE, "golden", name)) as fp:
        return json.load(fp)


def test_kat1_reference_notebook_vector():
    """examples/demo.ipynb:137-138 -- the only known-answer vector the reference holds for this path."""
    kat = _golden("kat1_ablaincourt.json")
    lx, ly = layout("Ablaincourt_")
    sol = floris_oracle.solve(lx, ly, kat["wind_speed"], kat["wind_direction"], kat["yaw"])
    # inputs and outputs are printed with 8 decimals: allow 2 units of the last printed digit
    assert np.max(np.abs(sol.ws_local - kat["local_wind_speed"])) < 2e-8
    assert np.max(np.abs(sol.wd_local - kat["local_wind_direction"])) < 2e-8


def test_kat2_floris_docs_example():
    kat = _golden("kat2_floris_docs.json")
    for case in kat["cases"]:
        sol = floris_oracle.solve(kat["layout_x"], kat["layout_y"], case["wind_speed"], 270.0, np.zeros(4))
        assert np.max(np.abs(sol.power_W / 1e3 - case["power_kW"])) < 1e-7


def test_self_consistency_vectors():
    """Outputs of the survey's independent restatement; regression, not reference-pinned."""
    vec = _golden("self_consistency.json")
    for case in vec["cases"]:
        if "layout" in case:
            lx, ly = layout(case["layout"])
        else:
            lx, ly = case["layout_x"], case["layout_y"]
        yaw = np.asarray(case["yaw"], dtype=np.float64) if "yaw" in case else np.float32(
            np.linspace(*case["yaw_linspace"])).astype(np.float64)
        sol = floris_oracle.solve(lx, ly, case["wind_speed"], case["wind_direction"], yaw)
        if "power_kW" in case:
            assert np.max(np.abs(sol.power_W / 1e3 - case["power_kW"])) < 1e-7, case
        if "farm_MW" in case:
            assert abs(sol.power_W.sum() / 1e6 - case["farm_MW"]) < 1e-9, case
        if "order" in case:
            assert list(sol.order) == case["order"]


def test_mean_of_nine_rule():
    """np.mean over the 9 identical grid x values equals fl(fl(8x+x)/9) -- the kernels rely on it."""
    rng = np.random.default_rng(0)
    x = rng.uniform(-1e4, 1e4, 5000)
    X = x[None, None, :, None, None] * np.ones((1, 1, x.size, 3, 3))
    m = np.array([np.mean(X[:, :, i:i + 1], axis=(3, 4))[0, 0, 0] for i in range(x.size)])
    assert np.array_equal(m, (8 * x + x) / 9)
    assert 0.01 < np.mean(m > x) < 0.06  # the self-mask is a real effect


def test_stable_sort_on_ties():
    lx, ly = layout("Turb6_Row2_")
    sol = floris_oracle.solve(lx, ly, 8.0, 270.0, np.zeros(6))
    assert list(sol.order) == [0, 3, 1, 4, 2, 5]


@pytest.mark.parametrize("name", CONFIG_LAYOUTS)
def test_c_oracle_matches_numpy_oracle(name):
    lx, ly = layout(name)
    T = len(lx)
    B = 3 if T > 40 else 6
    ws, wd = sample_winds(B, seed=11, tie_every=3)
    rng = np.random.default_rng(5)
    yaw = rng.uniform(-40, 40, (B, T)).astype(np.float32).astype(np.float64)
    c, s = host_trig(wd)
    got = c_oracle.solve_batch(lx, ly, ws, wd, yaw, cs=np.stack([c, s], 1))
    for b in range(B):
        ref = floris_oracle.solve(lx, ly, ws[b], wd[b], yaw[b])
        assert np.array_equal(got["order"][b], ref.order)
        for key, floor in (("power_W", 1.0), ("ws_local", 1e-3), ("wd_local", 1.0), ("ti", 1e-3), ("std_u", 1e-3),
                           ("std_v", 1e-3), ("std_w", 1e-3)):
            r = getattr(ref, key)
            assert np.max(np.abs(got[key][b] - r) / np.maximum(np.abs(r), floor)) < 1e-11, (name, b, key)


def test_zero_yaw_symmetry_and_wake_loss():
    """Domain sanity: a downstream turbine in a full wake produces less than the free-stream one; zero yaw, wd=270."""
    sol = floris_oracle.solve([0.0, 630.0, 1260.0], [0.0, 0.0, 0.0], 8.0, 270.0, np.zeros(3))
    assert sol.power_W[0] > sol.power_W[2] > 0 and sol.power_W[0] > sol.power_W[1] > 0
    assert np.all(sol.ti >= 0.06 - 1e-15)


def test_wake_steering_pays_off_on_an_aligned_row():
    """Physical plausibility of the yawed (parity-unpinned) branch: steering the wakes of an aligned row away from the
    downstream rotors raises the farm's total power, and the gain is roughly symmetric in the yaw sign (the asymmetry comes
    from the wake-rotation vortex of GCH)."""
    lx, ly = [0.0, 630.0, 1260.0], [0.0, 0.0, 0.0]
    base = floris_oracle.solve(lx, ly, 8.0, 270.0, [0.0, 0.0, 0.0]).power_W.sum()
    pos = floris_oracle.solve(lx, ly, 8.0, 270.0, [20.0, 10.0, 0.0]).power_W.sum()
    neg = floris_oracle.solve(lx, ly, 8.0, 270.0, [-20.0, -10.0, 0.0]).power_W.sum()
    assert pos > 1.10 * base and neg > 1.10 * base
    assert abs(pos - neg) / base < 0.05 and pos != neg
    # the yawed turbine itself loses power roughly like cos(yaw)^pP
    p0 = floris_oracle.solve(lx, ly, 8.0, 270.0, [0.0, 0.0, 0.0]).power_W[0]
    p20 = floris_oracle.solve(lx, ly, 8.0, 270.0, [20.0, 0.0, 0.0]).power_W[0]
    assert abs(p20 / p0 - np.cos(np.radians(20.0)) ** 1.88) < 0.01
