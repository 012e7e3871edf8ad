"""CPU tests: layout data (reference wfcrl/environments/data_cases.py) and the env-id grammar inputs."""
import numpy as np

from wfcrl_b200.layouts import get_layout, named_layouts


def test_turbine_counts_match_reference_code():
    counts = {k: get_layout(k)["num_turbines"] for k in named_layouts()}
    assert counts["HornsRev1_"] == 80   # README says 76; data_cases.py:269-291 has 80
    assert counts["HornsRev2_"] == 91 and counts["Ormonde_"] == 30 and counts["WMR_"] == 35
    assert counts["Turb_TCRWP_"] == 32 and counts["Turb32_Row5_"] == 32 and counts["Turb16_Row5_"] == 16
    assert counts["Ablaincourt_"] == 7 and counts["Turb6_Row2_"] == 6 and counts["Turb3_Row1_"] == 3
    for n in range(1, 13):
        assert counts[f"Turb{n}_Row1_"] == n


def test_procedural_rows_and_alias():
    row = get_layout("Turb5_Row1_")
    assert row["xcoords"] == [0.0, 504.0, 1008.0, 1512.0, 2016.0] and row["ycoords"] == [0.0] * 5
    assert row["dt"] == 60 and row["t_init"] == 0
    alias = get_layout("Turb16_TCRWP_")
    full = get_layout("Turb_TCRWP_")
    assert alias["num_turbines"] == 16 and alias["xcoords"] == full["xcoords"][:16]


def test_exact_x_ties_exist_at_270():
    lx = np.array(get_layout("Turb6_Row2_")["xcoords"])
    assert len(np.unique(lx)) == 3
