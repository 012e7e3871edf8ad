"""In-kernel auto-reset (wf_set_autoreset / wf_autoreset_finish): the truncating step resets the env's state itself and the
finish call (wind draw inside the geometry kernel + warm-up solve) restarts the episode.  It must reproduce, bit for bit, the
explicit chain `step -> wf_reset_sampled(truncated mask)` it replaces, and cost at most the geometry launch (+ vortex-table
build where a table exists), the warm-up launch (+ its FP64 re-solve launch on a strict FP32 handle)."""
import numpy as np
import pytest

from tests._util import layout, sample_winds

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_fused_autoreset_equals_explicit_reset_chain(cuda_device, precision):
    import torch

    from wfcrl_b200.backend import FlorisBatch

    lx, ly = layout("Turb16_TCRWP_")
    T, B, L = len(lx), 96, 6  # episodes of L steps: truncation at step L - 1 after the warm-up iteration
    ws, wd = sample_winds(B, seed=8)
    fused = FlorisBatch(lx, ly, B, precision=precision, kernel="fast", max_iter=L, reward_shaper="step")
    plain = FlorisBatch(lx, ly, B, precision=precision, kernel="fast", max_iter=L, reward_shaper="step")
    for fb in (fused, plain):
        fb.reset(ws, wd, host_trig=False)
    fused.set_autoreset(True, seed=77, env_id_offset=1000, turbulence_intensity_range=(0.05, 0.1))
    rng = np.random.default_rng(1)
    n_resets = 0
    for k in range(3 * L + 2):
        act = torch.as_tensor(rng.uniform(-5, 5, (B, T)).astype(np.float32), device="cuda")
        a = {key: v.clone() for key, v in fused.step(act).items()}
        b = {key: v.clone() for key, v in plain.step(act).items()}
        torch.cuda.synchronize()
        for key in a:
            assert torch.equal(a[key], b[key]), (k, key)
        if bool(b["truncated"].any()):
            assert bool(b["truncated"].all())
            n_resets += 1
            l0 = fused.launch_count()
            ra = {key: v.clone() for key, v in fused.autoreset_finish().items()}
            extra = fused.launch_count() - l0
            rb = {key: v.clone() for key, v in
                  plain.reset_sampled(b["truncated"].clone(), seed=77, env_id_offset=1000,
                                      turbulence_intensity_range=(0.05, 0.1)).items()}
            torch.cuda.synchronize()
            assert extra <= (4 if precision == "f32" else 3), extra  # geometry, table, warm-up (+ FP64 re-solve launch)
            for key in ("yaw", "wind_speed", "wind_direction", "freewind", "power", "load"):
                assert torch.equal(ra[key], rb[key]), (k, key)
            for name in ("ws", "wd", "ti_ambient", "episode", "num_iter", "num_moves", "yaw", "acc", "shaper_ref", "ws_norm"):
                assert np.array_equal(fused.get_state(name), plain.get_state(name)), (k, name)
        else:
            # finishing when nothing is marked changes nothing
            if k % 4 == 1:
                before = {key: v.clone() for key, v in fused.out.items()}
                fused.autoreset_finish()
                torch.cuda.synchronize()
                for key in before:
                    assert torch.equal(before[key], fused.out[key]), (k, key)
    assert n_resets == (3 * L + 2) // (L - 1)
    fused.close()
    plain.close()
