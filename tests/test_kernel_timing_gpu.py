"""wf_set_kernel_timing / wf_get_kernel_timing: the per-kernel device times bench.py's roofline block is built from."""
import numpy as np
import pytest

from tests._util import layout, sample_winds

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision,strict", [("f32", True), ("f32", False), ("f64", True)])
def test_kernel_timing_counts_calls_and_leaves_results_alone(cuda_device, precision, strict):
    import torch
    from wfcrl_b200.backend import FlorisBatch

    lx, ly = layout("Turb32_Row5_")
    T, B = len(lx), 512
    ws, wd = sample_winds(B, 3)
    ws[: B // 4] = np.linspace(3.0, 4.5, B // 4)  # low wind: the strict handle has envs to re-solve
    rng = np.random.default_rng(0)
    acts = [torch.as_tensor(rng.uniform(-5, 5, (B, T)).astype(np.float32), device="cuda") for _ in range(4)]
    outs = {}
    for timed in (False, True):
        fb = FlorisBatch(lx, ly, B, precision=precision, kernel="fast", max_iter=100, strict=strict)
        fb.reset(ws, wd, host_trig=True)
        if timed:
            fb.set_kernel_timing(True)
        for a in acts:
            o = fb.step(a)
        outs[timed] = {k: v.clone() for k, v in o.items()}
        if timed:
            t = fb.kernel_timing()
            assert t["calls"] == len(acts)
            assert 0.0 < t["step_kernel_ms"] < 50.0
            if precision == "f32" and strict:
                assert int(fb.get_state("ambiguous").sum()) > 0
                assert 0.0 < t["resolve_kernel_ms"] < 50.0
            else:
                assert t["resolve_kernel_ms"] < 0.05  # no second launch: just the gap between two events
            assert fb.kernel_timing()["calls"] == 0  # the query restarts the averages
            fb.set_kernel_timing(False)
            fb.step(acts[0])
            assert fb.kernel_timing()["calls"] == 0
        fb.close()
    for k in outs[False]:
        assert torch.equal(outs[False][k], outs[True][k]), k
