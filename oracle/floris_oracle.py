"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the FLORIS 3.5 steady-state wake solve that sits behind
every ``*_Floris`` env step of ifpen/wfcrl-env.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The product path
(``wfcrl_b200``) never imports anything from ``oracle/`` and fails loudly without its CUDA library.

What it restates
----------------
The arithmetic is NOT in the reference tree: it lives in the third-party dependency ``FLORIS==3.5``
(reference ``requirements.txt:8``), un-vendored and not installable in this container (no network,
no wheel).  The call sites that define what must be reproduced are

* ``wfcrl/interface.py:479``      ``tools.FlorisInterface(case.yaml)``
* ``wfcrl/interface.py:564``      ``fi.calculate_wake(yaw_angles=...)``
* ``wfcrl/interface.py:623``      ``fi.get_turbine_powers()``
* ``wfcrl/interface.py:632-636``  ``turbulence_intensity_field``, ``std(u|v|w, (3,4))``
* ``wfcrl/interface.py:643-647``  ``cbrt(mean(u**3))``, ``wd - degrees(arctan2(v, u))``
* ``wfcrl/interface.py:666``      ``fi.reinitialize(wind_speeds, wind_directions)``

with the model selected by ``wfcrl/simulators/floris/inputs/template/case.yaml:14-16,27-39,41-50,52-60,
76-89``: GCH = Gauss velocity deficit + Gauss deflection + Crespo-Hernandez wake-added turbulence + SOSFS
combination, secondary steering, yaw-added recovery and transverse velocities enabled, 3x3 rotor grid,
turbine ``nrel_5MW``.  The published FLORIS v3.5 algorithm (modules ``simulation/grid.py``,
``flow_field.py``, ``solver.py:sequential_solver``, ``turbine.py``, ``wake_deflection/gauss.py``,
``wake_velocity/gauss.py``, ``wake_turbulence/crespo_hernandez.py``, ``wake_combination/sosfs.py``,
``turbine_library/nrel_5MW.yaml``) is restated below following SURVEY.md Appendix A/B, keeping FLORIS'
(n_wd=1, n_ws=1, T, 3, 3) array shapes and numpy reduction calls so reduction orders are the ones numpy uses.

Parity status
-------------
* zero-yaw solve, local wind speed / direction: PINNED by the reference's own stored notebook output
  (``examples/demo.ipynb:137-138``, KAT-1; see ``tests/golden/kat1_ablaincourt.json``).
* per-turbine power: a recalled FLORIS-v3 documentation example (KAT-2; not a file in the reference tree) and, from
  the reference's own artefacts, KAT-3 below.
* yawed solves (down to the -40 deg bound), farm power, reward incl. the load-proxy term, single- and multi-agent
  transition logic: PINNED AT FIGURE RESOLUTION by KAT-3 (``tests/golden/kat3_notebook_curves.json``,
  ``tools/make_golden_curves.py``, ``tests/test_notebook_curves.py``): the farm-power curves FLORIS 3.5 produced for the
  two yawing episodes of ``examples/demo.ipynb`` (stored PNG outputs, digitised at 7.5e-4 / 4.3e-3 MW per pixel) and the
  printed episode totals (189.31593162, 192.22698147).  The winds of those episodes are not stored, so two numbers
  per episode are fitted; with them this oracle reproduces all 14 / 17 plateau levels within 0.27 / 0.23 pixel
  (2e-4 / 1e-3 MW, i.e. 2e-5 / 1e-4 of the farm power) AND the printed totals within 2.8e-6 / 4.2e-6 relative at the same
  time.  What stays unpinned beyond that resolution: per-turbine values of yawed solves and the individual load proxies
  (they enter the totals only as a 0.7 % term) -- only a live FLORIS 3.5 could pin those digit for digit.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------------------
# Model parameters (case.yaml) and turbine definition (FLORIS turbine_library/nrel_5MW.yaml)
# --------------------------------------------------------------------------------------------------
NUM_EPS = 0.001  # floris BaseModel.NUM_EPS

CASE = dict(
    air_density=1.225,          # case.yaml:31
    turbulence_intensity=0.06,  # case.yaml:33
    wind_shear=0.12,            # case.yaml:36
    wind_veer=0.0,              # case.yaml:39
    grid_points=3,              # case.yaml:16
    # wake_deflection_parameters.gauss, case.yaml:53-60
    ad=0.0, alpha=0.58, bd=0.0, beta=0.077, dm=1.0, ka=0.38, kb=0.004,
    # wake_turbulence_parameters.crespo_hernandez, case.yaml:85-89
    ch_initial=0.1, ch_constant=0.5, ch_ai=0.8, ch_downstream=-0.32,
)

NREL_5MW = dict(
    rotor_diameter=126.0, hub_height=90.0, TSR=8.0, pP=1.88, pT=1.88,
    generator_efficiency=1.0, ref_density_cp_ct=1.225, ref_tilt_cp_ct=5.0,
)

_WS_TABLE = np.array([0.0, 2.0, 2.5] + [3.0 + 0.5 * i for i in range(45)] + [25.01, 25.02, 50.0])
_CP_TABLE = np.array([
    0.0, 0.0, 0.0, 0.178085, 0.289075, 0.349022, 0.384728, 0.406059, 0.420228, 0.428823, 0.433873,
    0.436223, 0.436845, 0.436575, 0.436511, 0.436561, 0.436517, 0.435903, 0.434673, 0.433230, 0.430466, 0.378869,
    0.335199, 0.297991, 0.266092, 0.238588, 0.214748, 0.193981, 0.175808, 0.159835, 0.145741, 0.133256, 0.122157,
    0.112257, 0.103399, 0.095449, 0.088294, 0.081836, 0.075993, 0.070692, 0.065875, 0.061484, 0.057476, 0.053809,
    0.050447, 0.047358, 0.044518, 0.041900, 0.039483, 0.0, 0.0])
_CT_TABLE = np.array([
    0.0, 0.0, 0.0, 0.99, 0.99, 0.97373036, 0.92826162, 0.89210543, 0.86100905, 0.835423, 0.81237673,
    0.79225789, 0.77584769, 0.7629228, 0.76156073, 0.76261984, 0.76169723, 0.75232027, 0.74026851, 0.72987175,
    0.70701647, 0.54054532, 0.45509459, 0.39343381, 0.34250785, 0.30487242, 0.27164979, 0.24361964, 0.21973831,
    0.19918151, 0.18131868, 0.16537679, 0.15103727, 0.13998636, 0.1289037, 0.11970413, 0.11087113, 0.10339901,
    0.09617888, 0.09009926, 0.08395078, 0.0791188, 0.07448356, 0.07050731, 0.06684119, 0.06345518, 0.06032267,
    0.05741999, 0.05472609, 0.0, 0.0])
assert _WS_TABLE.shape == _CP_TABLE.shape == _CT_TABLE.shape == (51,)


def turbine_tables():
    """(wind_speed[51], thrust[51], inner_power[51]).  ``inner_power`` is what FLORIS' ``Turbine`` tabulates in
    ``__attrs_post_init__``: 0.5 * rotor_area * Cp * generator_efficiency * ws**3 (power / air density), and it is
    THIS table (not Cp) that ``power_interp`` interpolates linearly."""
    rotor_area = np.pi * (NREL_5MW["rotor_diameter"] / 2.0) ** 2.0
    inner_power = 0.5 * rotor_area * _CP_TABLE * NREL_5MW["generator_efficiency"] * _WS_TABLE ** 3
    return _WS_TABLE.copy(), _CT_TABLE.copy(), inner_power


def _interp_linear(x, xp, fp, left, right):
    """scipy ``interp1d(kind='linear', bounds_error=False, fill_value=(left, right))`` on sorted float64 1-D
    tables dispatches to ``np.interp`` for the in-range part; out-of-range values take the fill values."""
    x = np.asarray(x, dtype=np.float64)
    y = np.interp(x, xp, fp)
    y = np.where(x < xp[0], left, y)
    y = np.where(x > xp[-1], right, y)
    return y


def cosd(a):
    return np.cos(np.radians(a))


def sind(a):
    return np.sin(np.radians(a))


def wind_delta(wd):
    return ((wd - 270.0) % 360.0 + 360.0) % 360.0


# --------------------------------------------------------------------------------------------------
# Geometry (FLORIS grid.py TurbineGrid.set_grid + utilities.rotate_coordinates_rel_west) -- SURVEY A.2
# --------------------------------------------------------------------------------------------------
def rotate_layout(layout_x, layout_y, wd, cs=None):
    """Rotate the layout about the centre of its bounding box so the wind comes from the west.
    ``cs=(c, s)`` overrides cosd/sind of the deviation (used to feed a device-computed rotation back in)."""
    lx = np.asarray(layout_x, dtype=np.float64)
    ly = np.asarray(layout_y, dtype=np.float64)
    dev = wind_delta(np.float64(wd))
    c, s = (cosd(dev), sind(dev)) if cs is None else (np.float64(cs[0]), np.float64(cs[1]))
    xc = (np.min(lx) + np.max(lx)) / 2
    yc = (np.min(ly) + np.max(ly)) / 2
    xo = lx - xc
    yo = ly - yc
    xr = xo * c - yo * s + xc
    yr = xo * s + yo * c + yc
    return xr, yr


def sort_order(xr, order=None):
    """Canonical turbine order: STABLE ascending in rotated x (ties keep original index).  numpy's default
    argsort is only stable for T<=16; SURVEY section 7.3 makes the stable order canonical and allows a host override."""
    if order is not None:
        return np.asarray(order, dtype=np.int64)
    return np.argsort(xr, kind="stable")


class Solution:
    """Outputs of one solve, in the ORIGINAL (unsorted) turbine order, as WFCRL consumes them."""
    __slots__ = ("order", "u", "v", "w", "ti_field", "power_W", "ws_local", "wd_local", "ti", "std_u", "std_v",
                 "std_w", "self_mask", "x_sorted", "y_sorted")


def solve(layout_x, layout_y, ws, wd, yaw_deg, *, cs=None, order=None, ti_ambient=None):
    """One FLORIS GCH solve + the WFCRL measures derived from it (interface.py:557-577, 622-648).

    layout_x, layout_y : (T,) metres.   ws [m/s], wd [deg, already ``% 360`` -- interface.py:664].
    yaw_deg            : (T,) degrees in ORIGINAL turbine order (float32 values widened to float64, interface.py:562).
    """
    P = CASE
    D = NREL_5MW["rotor_diameter"]
    HH = NREL_5MW["hub_height"]
    TSR = NREL_5MW["TSR"]
    I0 = P["turbulence_intensity"] if ti_ambient is None else float(ti_ambient)
    shear = P["wind_shear"]
    veer = P["wind_veer"]
    G = P["grid_points"]
    T = len(layout_x)
    ws = np.float64(ws)
    wd = np.float64(wd)
    yaw = np.asarray(yaw_deg, dtype=np.float64).reshape(1, 1, T)

    # ---- A.2 geometry ------------------------------------------------------------------------------
    xr, yr = rotate_layout(layout_x, layout_y, wd, cs)
    zr = np.full(T, HH)
    disc_area_radius = 0.5 * D / 2
    disc_grid = np.linspace(-1 * disc_area_radius, disc_area_radius, G)
    template = np.ones((1, 1, T, G, G))
    _x = xr[None, None, :, None, None] * template
    _y = yr[None, None, :, None, None] + template * disc_grid[None, None, None, :, None]
    _z = zr[None, None, :, None, None] + template * disc_grid[None, None, None, None, :]
    srt = sort_order(xr, order)
    unsrt = np.argsort(srt, kind="stable")
    # FLORIS: sorted_indices = _x.argsort(axis=2) (5-D), np.take_along_axis(..., axis=2)
    srt5 = np.ascontiguousarray(np.broadcast_to(srt[None, None, :, None, None], (1, 1, T, G, G)))
    unsrt5 = np.ascontiguousarray(np.broadcast_to(unsrt[None, None, :, None, None], (1, 1, T, G, G)))
    x_s = np.take_along_axis(_x, srt5, axis=2)
    y_s = np.take_along_axis(_y, srt5, axis=2)
    z_s = np.take_along_axis(_z, srt5, axis=2)
    yaw_s = np.take_along_axis(yaw, srt[None, None, :], axis=2)

    # ---- A.3 initial flow --------------------------------------------------------------------------
    wind_profile_plane = (z_s / HH) ** shear
    dwind_profile_plane = shear * (1 / HH) ** shear * z_s ** (shear - 1)
    u_init = ws * wind_profile_plane
    dudz_init = ws * dwind_profile_plane
    u_s = u_init.copy()
    v_s = np.zeros_like(u_init)
    w_s = np.zeros_like(u_init)
    wake_field = np.zeros_like(u_init)
    turb_ti = I0 * np.ones((1, 1, T, 1, 1))
    Uinf = np.mean(u_init, axis=(2, 3, 4))[:, :, None, None, None]

    ws_tab, ct_tab, pw_tab = turbine_tables()
    eps = 0.2 * D
    self_mask = np.zeros(T, dtype=bool)

    # ---- A.4 .. A.8 sequential solver --------------------------------------------------------------
    for i in range(T):
        x_i = np.mean(x_s[:, :, i:i + 1], axis=(3, 4))[:, :, :, None, None]
        y_i = np.mean(y_s[:, :, i:i + 1], axis=(3, 4))[:, :, :, None, None]
        # z_i is computed by FLORIS but unused by the Gauss models (rCalt uses the hub height)
        u_i = u_s[:, :, i:i + 1]
        v_i = v_s[:, :, i:i + 1]
        yaw_i = yaw_s[:, :, i:i + 1, None, None]
        self_mask[i] = bool(x_s[0, 0, i, 0, 0] - x_i[0, 0, 0, 0, 0] < 0.0)

        # A.4  Ct, axial induction (turbine.py Ct / axial_induction; tilt factor cosd(5-5) = 1)
        avg_vel = np.cbrt(np.mean(u_i ** 3, axis=(3, 4)))  # (1,1,1)
        ct_raw = _interp_linear(avg_vel, ws_tab, ct_tab, 0.0001, 0.9999)
        ct_raw = np.clip(ct_raw, 0.0001, 0.9999)
        tilt_fac = cosd(NREL_5MW["ref_tilt_cp_ct"] - NREL_5MW["ref_tilt_cp_ct"])
        ct_i = (ct_raw * cosd(yaw_i[:, :, :, 0, 0]) * tilt_fac)[:, :, :, None, None]
        a_i = (0.5 / (cosd(yaw_i) * tilt_fac) * (1 - np.sqrt(1 - ct_i * cosd(yaw_i) * tilt_fac)))
        ti_i = turb_ti[:, :, i:i + 1]  # VIEW: sees the in-place yaw-added-recovery update below

        vel_top = ((HH + D / 2) / HH) ** shear
        vel_bottom = ((HH - D / 2) / HH) ** shear
        turbine_average_velocity = avg_vel[:, :, :, None, None]
        Gamma_wake_rotation = 0.25 * 2 * np.pi * D * (a_i - a_i ** 2) * turbine_average_velocity / TSR

        # A.5  secondary steering (wake_deflection/gauss.py wake_added_yaw)
        Gamma_top = (np.pi / 8) * D * vel_top * Uinf * ct_i
        Gamma_bottom = -1 * (np.pi / 8) * D * vel_bottom * Uinf * ct_i
        avg_v = np.mean(v_i, axis=(3, 4))
        yLocs = (y_s[:, :, i:i + 1] - y_i) + NUM_EPS
        z_own = z_s[:, :, i:i + 1]
        zT = z_own - (HH + D / 2) + NUM_EPS
        rT = yLocs ** 2 + zT ** 2
        core_shape = 1 - np.exp(-rT / (eps ** 2))
        v_top = np.mean((Gamma_top * zT) / (2 * np.pi * rT) * core_shape, axis=(3, 4))
        zB = z_own - (HH - D / 2) + NUM_EPS
        rB = yLocs ** 2 + zB ** 2
        core_shape = 1 - np.exp(-rB / (eps ** 2))
        v_bottom = np.mean((Gamma_bottom * zB) / (2 * np.pi * rB) * core_shape, axis=(3, 4))
        zC = z_own - HH + NUM_EPS
        rC = yLocs ** 2 + zC ** 2
        core_shape = 1 - np.exp(-rC / (eps ** 2))
        v_core = np.mean((Gamma_wake_rotation * zC) / (2 * np.pi * rC) * core_shape, axis=(3, 4))
        val = 2 * (avg_v - v_core) / (v_top + v_bottom)
        val = np.where(val < -1.0, -1.0, val)
        val = np.where(val > 1.0, 1.0, val)
        added_yaw = np.degrees(0.5 * np.arcsin(val))[:, :, :, None, None]
        eff_yaw_i = np.zeros_like(yaw_i)
        eff_yaw_i += yaw_i
        eff_yaw_i += added_yaw

        # A.6  Gauss deflection (wake_deflection/gauss.py GaussVelocityDeflection.function), opposite sign
        g = -1 * eff_yaw_i
        uR = u_init * ct_i * cosd(g) / (2.0 * (1 - np.sqrt(1 - (ct_i * cosd(g)))))
        u0 = u_init * np.sqrt(1 - ct_i)
        x0 = (D * (cosd(g) * (1 + np.sqrt(1 - ct_i * cosd(g))))
              / (np.sqrt(2) * (4 * P["alpha"] * ti_i + 2 * P["beta"] * (1 - np.sqrt(1 - ct_i)))) + x_i)
        ky = P["ka"] * ti_i + P["kb"]
        kz = P["ka"] * ti_i + P["kb"]
        C0 = 1 - u0 / u_init
        M0 = C0 * (2 - C0)
        E0 = C0 ** 2 - 3 * np.exp(1.0 / 12.0) * C0 + 3 * np.exp(1.0 / 3.0)
        sigma_z0 = D * 0.5 * np.sqrt(uR / (u_init + u0))
        sigma_y0 = sigma_z0 * cosd(g) * cosd(veer)
        xR = x_i
        theta_c0 = P["dm"] * (0.3 * np.radians(g) / cosd(g))
        theta_c0 = theta_c0 * (1 - np.sqrt(1 - ct_i * cosd(g)))
        delta0 = np.tan(theta_c0) * (x0 - x_i)
        delta_near = ((x_s - xR) / (x0 - xR)) * delta0 + (P["ad"] + P["bd"] * (x_s - x_i))
        delta_near = delta_near * np.array(x_s >= xR)
        delta_near = delta_near * np.array(x_s <= x0)
        sigma_y = ky * (x_s - x0) + sigma_y0
        sigma_z = kz * (x_s - x0) + sigma_z0
        sigma_y = sigma_y * np.array(x_s >= x0) + sigma_y0 * np.array(x_s < x0)
        sigma_z = sigma_z * np.array(x_s >= x0) + sigma_z0 * np.array(x_s < x0)
        ln_num = (1.6 + np.sqrt(M0)) * (1.6 * np.sqrt(sigma_y * sigma_z / (sigma_y0 * sigma_z0)) - np.sqrt(M0))
        ln_den = (1.6 - np.sqrt(M0)) * (1.6 * np.sqrt(sigma_y * sigma_z / (sigma_y0 * sigma_z0)) + np.sqrt(M0))
        delta_far = (delta0 + theta_c0 * E0 / 5.2 * np.sqrt(sigma_y0 * sigma_z0 / (ky * kz * M0))
                     * np.log(ln_num / ln_den) + (P["ad"] + P["bd"] * (x_s - x_i)))
        delta_far = delta_far * np.array(x_s > x0)
        deflection = delta_near + delta_far

        # A.7  transverse velocities (calculate_transverse_velocity) -- uses yaw_i, NOT the effective yaw
        delta_x = x_s - x_i
        yL = (y_s - y_i) + NUM_EPS
        Gt = sind(yaw_i) * cosd(yaw_i) * ((np.pi / 8) * D * vel_top * Uinf * ct_i)
        Gb = -1 * sind(yaw_i) * cosd(yaw_i) * ((np.pi / 8) * D * vel_bottom * Uinf * ct_i)
        lmda = D / 8
        kappa = 0.41
        lm = kappa * z_s / (1 + kappa * z_s / lmda)
        nu = lm ** 2 * np.abs(dudz_init)
        decay = eps ** 2 / (4 * nu * delta_x / Uinf + eps ** 2)
        zT = z_s - (HH + D / 2) + NUM_EPS
        rT = yL ** 2 + zT ** 2
        core_shape = 1 - np.exp(-rT / (eps ** 2))
        V1 = (Gt * zT) / (2 * np.pi * rT) * core_shape * decay
        W1 = (-1 * Gt * yL) / (2 * np.pi * rT) * core_shape * decay
        zB = z_s - (HH - D / 2) + NUM_EPS
        rB = yL ** 2 + zB ** 2
        core_shape = 1 - np.exp(-rB / (eps ** 2))
        V2 = (Gb * zB) / (2 * np.pi * rB) * core_shape * decay
        W2 = (-1 * Gb * yL) / (2 * np.pi * rB) * core_shape * decay
        zC = z_s - HH + NUM_EPS
        rC = yL ** 2 + zC ** 2
        core_shape = 1 - np.exp(-rC / (eps ** 2))
        V5 = (Gamma_wake_rotation * zC) / (2 * np.pi * rC) * core_shape * decay
        W5 = (-1 * Gamma_wake_rotation * yL) / (2 * np.pi * rC) * core_shape * decay
        zTb = z_s + (HH + D / 2) + NUM_EPS
        rTb = yL ** 2 + zTb ** 2
        core_shape = 1 - np.exp(-rTb / (eps ** 2))
        V3 = (-1 * Gt * zTb) / (2 * np.pi * rTb) * core_shape * decay
        W3 = (Gt * yL) / (2 * np.pi * rTb) * core_shape * decay
        zBb = z_s + (HH - D / 2) + NUM_EPS
        rBb = yL ** 2 + zBb ** 2
        core_shape = 1 - np.exp(-rBb / (eps ** 2))
        V4 = (-1 * Gb * zBb) / (2 * np.pi * rBb) * core_shape * decay
        W4 = (Gb * yL) / (2 * np.pi * rBb) * core_shape * decay
        zCb = z_s + HH + NUM_EPS
        rCb = yL ** 2 + zCb ** 2
        core_shape = 1 - np.exp(-rCb / (eps ** 2))
        V6 = (-1 * Gamma_wake_rotation * zCb) / (2 * np.pi * rCb) * core_shape * decay
        W6 = (Gamma_wake_rotation * yL) / (2 * np.pi * rCb) * core_shape * decay
        v_wake = V1 + V2 + V3 + V4 + V5 + V6
        w_wake = W1 + W2 + W3 + W4 + W5 + W6
        v_wake[delta_x < 0.0] = 0.0
        w_wake[delta_x < 0.0] = 0.0
        w_wake[w_wake < 0.0] = 0.0

        # yaw-added recovery (yaw_added_turbulence_mixing); TI taken at grid point (0,0) of the source
        I_i = ti_i[:, :, 0, 0, 0]
        average_u_i = np.cbrt(np.mean(u_i ** 3, axis=(2, 3, 4)))
        k = (average_u_i * I_i) ** 2 / (2 / 3)
        u_term = np.sqrt(2 * k)
        v_term = np.mean(v_i + v_wake[:, :, i:i + 1], axis=(2, 3, 4))
        w_term = np.mean(w_s[:, :, i:i + 1] + w_wake[:, :, i:i + 1], axis=(2, 3, 4))
        k_total = 0.5 * (u_term ** 2 + v_term ** 2 + w_term ** 2)
        I_total = np.sqrt((2 / 3) * k_total) / average_u_i
        I_mixing = (I_total - I_i)[:, :, None, None, None]
        turb_ti[:, :, i:i + 1] = ti_i + 2 * I_mixing  # in place => ti_i (a view) is updated too

        # A.8  Gauss velocity deficit (wake_velocity/gauss.py), opposite sign, uses yaw_i
        gv = -1 * yaw_i
        uR = u_init * ct_i / (2.0 * (1 - np.sqrt(1 - ct_i)))
        u0 = u_init * np.sqrt(1 - ct_i)
        sigma_z0 = D * 0.5 * np.sqrt(uR / (u_init + u0))
        sigma_y0 = sigma_z0 * cosd(gv) * cosd(veer)
        xR = x_i
        x0 = np.ones_like(u_init)
        x0 = x0 * (D * cosd(gv) * (1 + np.sqrt(1 - ct_i)))
        x0 = x0 / (np.sqrt(2) * (4 * P["alpha"] * ti_i + 2 * P["beta"] * (1 - np.sqrt(1 - ct_i))))
        x0 = x0 + x_i
        velocity_deficit = np.zeros_like(u_init)
        near_mask = np.array(x_s > xR + 0.1) * np.array(x_s < x0)
        far_mask = np.array(x_s >= x0)

        def r_c(sig_y, sig_z):
            vr = np.deg2rad(veer)
            a = np.cos(vr) ** 2 / (2 * sig_y ** 2) + np.sin(vr) ** 2 / (2 * sig_z ** 2)
            b = -np.sin(2 * vr) / (4 * sig_y ** 2) + np.sin(2 * vr) / (4 * sig_z ** 2)
            c = np.sin(vr) ** 2 / (2 * sig_y ** 2) + np.cos(vr) ** 2 / (2 * sig_z ** 2)
            dy = y_s - y_i - deflection
            dz = z_s - HH
            r = a * (dy ** 2) - 2 * b * dy * dz + c * (dz ** 2)
            d = np.clip(1 - (ct_i * cosd(gv) / (8.0 * sig_y * sig_z / (D * D))), 0.0, 1.0)
            return r, 1 - np.sqrt(d)

        if np.sum(near_mask):
            ramp_up = (x_s - xR) / (x0 - xR)
            ramp_down = (x0 - x_s) / (x0 - xR)
            sig_y = ramp_down * 0.501 * D * np.sqrt(ct_i / 2.0) + ramp_up * sigma_y0
            sig_y = sig_y * np.array(x_s >= xR) + np.ones_like(sig_y) * np.array(x_s < xR) * 0.5 * D
            sig_z = ramp_down * 0.501 * D * np.sqrt(ct_i / 2.0) + ramp_up * sigma_z0
            sig_z = sig_z * np.array(x_s >= xR) + np.ones_like(sig_z) * np.array(x_s < xR) * 0.5 * D
            r, C = r_c(sig_y, sig_z)
            near_def = C * np.exp(-1 * r ** 1 / (2 * np.sqrt(0.5) ** 2))
            velocity_deficit = velocity_deficit + near_def * near_mask
        if np.sum(far_mask):
            kyv = P["ka"] * ti_i + P["kb"]
            kzv = P["ka"] * ti_i + P["kb"]
            sig_y = (kyv * (x_s - x0) + sigma_y0) * far_mask + sigma_y0 * np.array(x_s < x0)
            sig_z = (kzv * (x_s - x0) + sigma_z0) * far_mask + sigma_z0 * np.array(x_s < x0)
            r, C = r_c(sig_y, sig_z)
            far_def = C * np.exp(-1 * r ** 1 / (2 * np.sqrt(0.5) ** 2))
            velocity_deficit = velocity_deficit + far_def * far_mask

        # SOSFS combination
        wake_field = np.hypot(wake_field, velocity_deficit * u_init)

        # Crespo-Hernandez wake-added turbulence
        ch_dx = x_s - x_i
        up_mask = np.array(ch_dx <= 0.1, dtype=float)
        dn_mask = np.array(ch_dx > -0.1, dtype=float)
        ch_dx = ch_dx * dn_mask + np.ones_like(ch_dx) * up_mask
        with np.errstate(divide="ignore", invalid="ignore"):
            wat = (P["ch_constant"] * a_i ** P["ch_ai"] * I0 ** P["ch_initial"]
                   * (ch_dx / D) ** P["ch_downstream"])
        wat = wat * dn_mask
        area_overlap = np.sum(velocity_deficit * u_init > 0.05, axis=(3, 4)) / (G * G)
        area_overlap = area_overlap[:, :, :, None, None]
        ti_added = (area_overlap * np.nan_to_num(wat, posinf=0.0) * np.array(x_s > x_i)
                    * np.array(np.abs(y_i - y_s) < 2 * D) * np.array(x_s <= 15 * D + x_i))
        turb_ti = np.maximum(np.sqrt(ti_added ** 2 + I0 ** 2), turb_ti)

        u_s = u_init - wake_field
        v_s = v_s + v_wake
        w_s = w_s + w_wake

    # ---- A.9 finalise (unsort) ---------------------------------------------------------------------
    turb_ti = turb_ti * np.ones((1, 1, T, G, G))
    u = np.take_along_axis(u_s, unsrt5, axis=2)
    v = np.take_along_axis(v_s, unsrt5, axis=2)
    w = np.take_along_axis(w_s, unsrt5, axis=2)
    ti_field = np.take_along_axis(turb_ti, unsrt5, axis=2)
    ti_avg = np.mean(ti_field, axis=(3, 4))

    # ---- A.10 power (turbine.py rotor_effective_velocity / power) and WFCRL measures ------------------
    avg_vel_all = np.cbrt(np.mean(u ** 3, axis=(3, 4)))
    pW = NREL_5MW["pP"] / 3.0
    pV = NREL_5MW["pT"] / 3.0
    rho = P["air_density"]
    ref_rho = NREL_5MW["ref_density_cp_ct"]
    veff = ((rho / ref_rho) ** (1 / 3) * avg_vel_all * cosd(yaw) ** pW * cosd(0.0) ** pV)
    p = _interp_linear(veff, ws_tab, pw_tab, 0.0, 0.0)
    power_W = (p * ref_rho).flatten()

    sol = Solution()
    sol.order = srt.copy()
    sol.u, sol.v, sol.w = u[0, 0], v[0, 0], w[0, 0]
    sol.ti_field = ti_field[0, 0]
    sol.power_W = power_W
    sol.ws_local = avg_vel_all.squeeze().reshape(T)                                   # interface.py:643
    sol.wd_local = np.mean(wd - np.degrees(np.arctan2(v, u)), axis=(3, 4)).squeeze().reshape(T)  # :644-647
    sol.ti = ti_avg.squeeze().reshape(T)                                              # :632
    sol.std_u = np.std(u, (3, 4)).squeeze().reshape(T)                                # :634
    sol.std_v = np.std(v, (3, 4)).squeeze().reshape(T)
    sol.std_w = np.std(w, (3, 4)).squeeze().reshape(T)
    sol.self_mask = self_mask
    sol.x_sorted = xr[srt].copy()
    sol.y_sorted = yr[srt].copy()
    return sol
