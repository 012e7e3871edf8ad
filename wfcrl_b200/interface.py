"""Simulator interface: the drop-in boundary of the reference (wfcrl/interface.py:25-50 ``BaseInterface`` and the
de-facto protocol ``WindFarmMDP`` uses).

``FlorisInterface`` here has the reference class's name, constructor arguments, attributes and method contracts
(wfcrl/interface.py:444-671) but no FLORIS inside: every ``update_command`` is one launch of the sm_100a step kernel in
FP64 interface mode (warp-per-env kernel by default, ``kernel="basic"`` selects the one-thread-per-turbine kernel)
through the C-ABI (``wf_update_command``), as a batch of one environment.  The batched environments
(``wfcrl_b200.vector_env``) use the same kernels in env mode without this per-call host round trip.
"""
from __future__ import annotations

import itertools
import time
import warnings
from abc import ABC
from typing import List, Union

import numpy as np

from .environments.data_cases import FarmCase


def _load_series(time_series) -> np.ndarray:
    """Rows [speed, direction, ...] of a wind time series given as an array or as the path of a csv file with a header row."""
    if isinstance(time_series, str):
        import pandas as pd

        time_series = pd.read_csv(time_series).values
    rows = np.asarray(time_series)
    if rows.ndim != 2 or rows.shape[1] < 2:
        raise AssertionError("a wind time series holds rows of [speed, direction]")
    return rows


class BaseInterface(ABC):
    def __init__(self):
        self.num_turbines = None

    @property
    def wind_speed(self):
        pass

    @property
    def wind_dir(self):
        pass

    def set_yaw_angles(self, yaws: List):
        pass

    def get_yaw_angles(self) -> List:
        pass

    def avg_powers(self) -> List:
        pass

    def init(self):
        pass

    def next_wind(self):
        pass


class FlorisInterface(BaseInterface):
    CONTROL_SET = ["yaw"]
    DEFAULT_MEASURE_MAP = {
        "yaw": 0,
        "wind_speed": 1,
        "wind_direction": 2,
        "load": [3, 4, 5, 6],
        "freewind_measurements": None,
    }

    def __init__(self, num_turbines: int, simul_file=None, max_iter: int = int(1e4), log_file: str = None,
                 wind_speed: float = None, wind_direction: float = None,
                 wind_time_series: Union[str, np.ndarray] = None, *, xcoords=None, ycoords=None, device: int = 0,
                 precision: str = "f64", kernel: str = "fast"):
        """``simul_file``: path of a FLORIS v3 input yaml as in the reference (layout, flow field and Gauss / GCH /
        Crespo-Hernandez parameters are read from it; models the kernels do not implement are refused, see
        ``floris_yaml.py``), or a dict with ``xcoords``/``ycoords``, or None when the coordinates are passed by keyword
        (flow / wake parameters then are the template's, wfcrl/simulators/floris/inputs/template/case.yaml)."""
        super().__init__()
        import torch

        from .backend import FlorisBatch

        overrides = None
        if isinstance(simul_file, str):
            from .floris_yaml import load_floris_yaml

            parsed = load_floris_yaml(simul_file)
            xcoords, ycoords, overrides = parsed["xcoords"], parsed["ycoords"], parsed["overrides"]
            wind_speed = parsed["wind_speed"] if wind_speed is None else wind_speed
            wind_direction = parsed["wind_direction"] if wind_direction is None else wind_direction
        elif isinstance(simul_file, dict):
            xcoords, ycoords = simul_file["xcoords"], simul_file["ycoords"]
        if xcoords is None or ycoords is None:
            raise ValueError("FlorisInterface needs the turbine coordinates (xcoords, ycoords)")
        assert len(xcoords) == num_turbines == len(ycoords)
        self.num_turbines = num_turbines
        self._torch = torch
        if overrides and overrides.get("turbine_grid_points", 3) != 3:
            kernel = "basic"  # the tuned kernels are built for the template's 3x3 rotor grid
        self.fi = FlorisBatch(xcoords, ycoords, 1, device=device, precision=precision, kernel=kernel,
                              max_iter=int(max_iter), config_overrides=overrides)
        self.measure_map = self.DEFAULT_MEASURE_MAP
        self._num_measures = 7
        self.dt = 60
        self.max_iter = max_iter
        self._logging = False
        self._wind_speed, self._wind_dir = 8.0, 270.0
        self.wind_time_series = wind_time_series
        self.wind_generator = self._make_wind_generator(wind_speed, wind_direction, wind_time_series)
        wind_speed, wind_direction = next(self.wind_generator)
        self.init(wind_speed, wind_direction)
        if log_file is not None:
            self._log_file = log_file
            self._logging = True

    # -- construction -------------------------------------------------------------------------------------------
    @classmethod
    def from_case(cls, case: FarmCase, log_file: str = None, output_dir: str = None):
        params = case.simul_params
        return cls(num_turbines=case.num_turbines, simul_file=None, max_iter=case.max_iter, log_file=log_file,
                   wind_speed=float(params["speed"]), wind_direction=float(params["direction"]),
                   wind_time_series=params["wind_time_series"], xcoords=params["xcoords"], ycoords=params["ycoords"])

    def _make_wind_generator(self, wind_speed=None, wind_direction=None, time_series=None):
        """Iterator of (speed, direction) pairs feeding ``update_wind`` before every solve (reference behaviour,
        wfcrl/interface.py:503-524): a steady wind repeats forever; a time series (ndarray or csv path, columns speed,
        direction) is played ONCE, starting from a row drawn with numpy's GLOBAL generator and wrapping around to the row
        before it -- so ``np.random.seed`` controls the start, and a series shorter than the episode ends in StopIteration,
        exactly like the reference."""
        if time_series is None:
            return itertools.repeat((wind_speed, wind_direction))
        rows = _load_series(time_series)
        first = np.random.randint(0, rows.shape[0])
        return iter(np.roll(rows, -first, axis=0))

    # -- wind -----------------------------------------------------------------------------------------------------
    @property
    def wind_speed(self):
        return self._wind_speed

    @property
    def wind_dir(self):
        return self._wind_dir

    def update_wind(self, wind_speed: float = None, wind_direction: float = None):
        wind_direction = wind_direction % 360
        if wind_speed != self._wind_speed or wind_direction != self._wind_dir or not self._wind_pushed:
            self._wind_speed, self._wind_dir = float(wind_speed), float(wind_direction)
            torch = self._torch
            dev = self.fi.device
            self.fi.update_wind(torch.tensor([self._wind_speed], dtype=torch.float64, device=dev),
                                torch.tensor([self._wind_dir], dtype=torch.float64, device=dev), host_trig=True)
            self._wind_pushed = True

    def init(self, wind_speed: float = None, wind_direction: float = None):
        has_series = self.wind_time_series is not None and not (
            isinstance(self.wind_time_series, str) and not self.wind_time_series)
        if has_series and wind_speed is not None:
            warnings.warn(f"Wind speed = {wind_speed} requested, but wind_time_series mode is activated. "
                          "Request will be ignored.")
            wind_speed = None
        if has_series and wind_direction is not None:
            warnings.warn(f"Wind direction = {wind_direction} requested, but wind_time_series mode is activated. "
                          "Request will be ignored.")
            wind_direction = None
        self.wind_generator = self._make_wind_generator(wind_speed, wind_direction,
                                                        self.wind_time_series if has_series else None)
        ws, wd = next(self.wind_generator)
        # device-side reset of counters / yaw command, then push the wind (geometry pass)
        self.fi.reset(ws, wd % 360, host_trig=True, warmup_solves=0)
        self._wind_speed, self._wind_dir = float(ws), float(wd % 360)
        self._wind_pushed = True
        self._num_iter = 0
        self._current_yaw_command = np.zeros((1, 1, self.num_turbines))
        self.current_measures = np.zeros((self.num_turbines, self._num_measures)) * np.nan
        self._powers = np.zeros(self.num_turbines)

    # -- the hot call ---------------------------------------------------------------------------------------------
    def update_command(self, yaw: np.ndarray = None):
        if yaw is not None:
            self._current_yaw_command[0, 0, :] = np.asarray(yaw).astype(np.double)
        self.update_wind(*next(self.wind_generator))
        out = self.fi.update_command_host(self._current_yaw_command.reshape(1, -1))  # one library call, host buffers
        self.current_measures[:, self.measure_map["yaw"]] = self._current_yaw_command[0, 0]
        self.current_measures[:, self.measure_map["wind_speed"]] = out["wind_speed"][0]
        self.current_measures[:, self.measure_map["wind_direction"]] = out["wind_direction"][0]
        self.current_measures[:, self.measure_map["load"]] = out["load"][0]  # already x1e7
        self._powers = out["power"][0].astype(np.float64)  # W in interface mode
        self._num_iter += 1
        if self._logging:
            self._append_log()
        return self._num_iter == self.max_iter

    def _append_log(self):
        """One line per solve in the reference's log format (wfcrl/interface.py:579-585)."""
        fields = (("Sent command YAW", self.get_yaw_command()), ("***********Received Power:", self.avg_powers()),
                  ("Wind :", self.avg_wind()))
        line = f"{fields[0][0]} {fields[0][1]} - {fields[1][0]} {fields[1][1]} {fields[2][0]} {fields[2][1]}\n"
        with open(self._log_file, "a") as fp:
            fp.write(line)

    # -- measures -------------------------------------------------------------------------------------------------
    def get_yaw_command(self):
        return self._current_yaw_command.copy().flatten()

    def avg_farm_power(self):
        return self.avg_powers().sum()

    def avg_powers(self) -> np.ndarray:
        return self._powers.copy()

    def avg_wind(self) -> np.ndarray:
        return np.array([self.wind_speed, self.wind_dir]).squeeze()

    def get_measure(self, measure: str):
        if measure not in self.measure_map:
            return None
        if measure == "freewind_measurements":
            return self.avg_wind()
        return self.current_measures[:, self.measure_map[measure]].copy()

    def get_parameters(self):
        pass

    def sample_parameters(self):
        pass

    def __repr__(self):
        return f"<wfcrl_b200.FlorisInterface {self.num_turbines} turbines, created {time.strftime('%H:%M:%S')}>"
