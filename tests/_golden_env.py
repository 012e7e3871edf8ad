"""Replay of tests/golden/env_ref_*.json: step records the UNMODIFIED reference env classes produced in the build container
(tools/make_golden_env.py; FLORIS replaced by the numpy oracle) -- shared by the CPU test of oracle/env_oracle.py and the
GPU tests of the drop-in envs and of the batched kernels."""
import glob
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SINGLE = sorted(os.path.basename(p)[8:-5] for p in glob.glob(os.path.join(HERE, "golden", "env_ref_single_*.json")))
MULTI = sorted(os.path.basename(p)[8:-5] for p in glob.glob(os.path.join(HERE, "golden", "env_ref_multi_*.json")))


def load(name):
    with open(os.path.join(HERE, "golden", f"env_ref_{name}.json")) as fp:
        return json.load(fp)


def arr(rec):
    return np.array(rec["data"], dtype=rec["dtype"]).reshape(rec["shape"])


def layout_name(env_id):
    name = env_id[4:] if env_id.startswith("Dec_") else env_id
    return name[: -len("Floris")]


def shaper_spec(make_kwargs):
    """("none" | "reference" | "step", reference value) from the string the generator stored."""
    text = make_kwargs.get("reward_shaper")
    if text is None:
        return "none", 0.0
    kind, arg = text.rstrip(")").split("(")
    return {"ReferencePercentage": "reference", "StepPercentage": "step"}[kind], float(arg) if arg else 0.0


def compare_obs(got, ref, tol, where):
    for key, rec in ref.items():
        want = arr(rec)
        have = np.asarray(got[key])
        assert have.shape == want.shape, (where, key, have.shape, want.shape)
        if key == "yaw":  # float32 state: bit-exact
            assert np.array_equal(have.astype(np.float64), want.astype(np.float64)), (where, key, have, want)
        else:
            assert np.allclose(have, want, rtol=tol, atol=tol * 1e-3), (where, key, np.max(np.abs(have - want)))


def close(a, b, tol, floor=0.0):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return bool(np.all(np.abs(a - b) <= tol * np.maximum(np.abs(b), floor) + 1e-300))
