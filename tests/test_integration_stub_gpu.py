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
