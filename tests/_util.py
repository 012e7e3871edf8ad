"""Shared helpers for the parity tests (inputs follow SURVEY.md section 8d)."""
import numpy as np

from wfcrl_b200.layouts import get_layout

CONFIG_LAYOUTS = ["Turb6_Row2_", "Ablaincourt_", "Turb_TCRWP_", "Turb16_TCRWP_", "Turb32_Row5_", "HornsRev1_"]


def sample_winds(n, seed=0, tie_every=0):
    """ws = clip(8*Weibull(8), 3, 28), wd = N(270, 20) % 360: the reset distribution of wfcrl/mdp.py:242-258."""
    rng = np.random.default_rng(seed)
    ws = np.clip(8 * rng.weibull(8, n), 3, 28)
    wd = np.clip(rng.normal(270, 20, n) % 360, 0, 360)
    if tie_every:
        wd[::tie_every] = 270.0  # exact x-ties for the row layouts
    return ws, wd


def host_trig(wd):
    dev = (((np.asarray(wd) % 360.0) - 270.0) % 360.0 + 360.0) % 360.0
    return np.cos(np.radians(dev)), np.sin(np.radians(dev))


def layout(name):
    c = get_layout(name)
    return np.asarray(c["xcoords"], dtype=np.float64), np.asarray(c["ycoords"], dtype=np.float64)


def rel_err(a, ref, floor):
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return np.max(np.abs(a - ref) / np.maximum(np.abs(ref), floor))
