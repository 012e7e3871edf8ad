"""The reference-side ctypes binding shown in INTEGRATION.md is executed verbatim (only the library path is rewritten)
and checked against the oracle: the documentation cannot rot."""
import os
import re

import numpy as np
import pytest

from oracle import c_oracle
from tests._util import layout

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_integration_md_stub_runs_and_matches_oracle(cuda_device):
    from wfcrl_b200 import _lib
    from wfcrl_b200.environments.data_cases import floris_case
    from wfcrl_b200.interface import BaseInterface, FlorisInterface

    _lib.load()
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(.*?)```", text, flags=re.S).group(1)
    code = code.replace('C.CDLL("libwfcrl_b200.so")', f'C.CDLL({_lib.library_path()!r})')
    ns = {"BaseInterface": BaseInterface, "FlorisInterface": FlorisInterface}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    case = floris_case("Ablaincourt_")
    case.max_iter = 4
    iface = ns["B200FlorisInterface"].from_case(case)
    iface.init(6.48958384, 266.363907)
    assert iface.update_command() is False
    ws_l = iface.get_measure("wind_speed")
    assert np.max(np.abs(ws_l - [6.46819497, 4.58929161, 6.46702757, 6.21243961, 6.20072934, 6.1100638, 5.76785291])) < 2e-8
    yaw = np.array([10, 0, -5, 0, 20, 0, 0], dtype=np.float32)
    assert iface.update_command(yaw=yaw) is False
    lx, ly = layout("Ablaincourt_")
    ref = c_oracle.solve(lx, ly, 6.48958384, 266.363907, yaw.astype(np.float64))
    assert np.max(np.abs(iface.avg_powers() - ref.power_W) / ref.power_W) < 1e-9
    assert np.allclose(iface.get_measure("load")[:, 0], ref.ti * 1e7, rtol=1e-9)
    assert iface.get_measure("pitch") is None and np.allclose(iface.get_measure("freewind_measurements"), [6.48958384, 266.363907])
    iface.update_command()
    assert iface.update_command() is True  # 4th iteration == max_iter


def test_plain_c_client_of_the_abi(cuda_device, tmp_path):
    """examples/c_abi_example.c: a C program with no Python / torch in the process drives the library through the
    host-buffer entry points and prints the reference notebook's reset observation (KAT-1) and one env step."""
    import subprocess

    from wfcrl_b200 import _lib

    libdir = os.path.dirname(_lib.library_path())
    exe = str(tmp_path / "c_abi_example")
    subprocess.run(["gcc", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_example.c"),
                    "-L" + libdir, "-lwfcrl_b200", "-Wl,-rpath," + libdir, "-lm", "-o", exe], check=True)
    text = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    rows = {line.split(":")[0]: line.split(":", 1)[1] for line in text.strip().splitlines()}
    ws_l = np.array(rows["local wind speed"].split(), dtype=float)
    wd_l = np.array(rows["local wind direction"].split(), dtype=float)
    assert np.max(np.abs(ws_l - [6.46819497, 4.58929161, 6.46702757, 6.21243961, 6.20072934, 6.1100638, 5.76785291])) < 2e-8
    assert np.max(np.abs(wd_l - [266.64538262, 267.04575667, 266.77944635, 266.86120544, 266.89071421, 266.92108378,
                                 266.99680007])) < 2e-8
    assert rows["yaw after step"].split() == ["-5.0"] + ["0.0"] * 6
    lx, ly = layout("Ablaincourt_")
    ref = c_oracle.solve(lx, ly, 6.48958384, 266.363907, np.array([-5.0, 0, 0, 0, 0, 0, 0]))
    power = np.array(rows["power MW"].split(), dtype=float)
    assert np.max(np.abs(power - ref.power_W / 1e6) / (ref.power_W / 1e6)) < 1e-8      # printed with 9 decimals
    loads = np.stack([ref.ti, ref.std_u, ref.std_v, ref.std_w], 1)
    reward = np.mean(ref.power_W / 1e6 * 1e3 / 6.48958384 ** 3) - 0.1 * np.mean(np.abs(loads))
    got = float(rows["reward"].split()[0])
    assert abs(got - reward) < 1e-9 * abs(reward) and "truncated: 0" in text
