"""Single-environment MDP over a simulator interface (host side of the drop-in path).

Same public surface as the reference's ``WindFarmMDP`` (wfcrl/mdp.py:19-319): ``action_space`` / ``state_space``
(Dict of Box), ``reset(seed, options)``, ``take_action``, ``step_interface``, ``get_controlled_state_transition``,
``get_state_powers``, ``get_accumulated_actions`` and the class constants.  The wake solve behind
``interface.update_command`` runs on the GPU; the arithmetic kept here is the float32 yaw bookkeeping of a single env,
written with the same numpy dtypes as the reference so a batch of one reproduces it bit for bit.  The batched path
(``wfcrl_b200.vector_env``) fuses all of this into the step kernel instead.
"""
from __future__ import annotations

import copy
from collections import OrderedDict
from collections.abc import Iterable
from typing import Dict, Type, Union
from warnings import warn

import numpy as np

from . import spaces
from .environments.data_cases import FarmCase
from .interface import BaseInterface


def _action_space(controls: dict, num_turbines: int, continuous: bool):
    """Per-control action space: a symmetric Box of half-width ``step`` (continuous) or {down, hold, up} per turbine."""
    if continuous:
        return spaces.Dict({name: spaces.Box(-spec[2], spec[2], shape=(num_turbines,))
                            for name, spec in controls.items()})
    return spaces.Dict({name: spaces.MultiDiscrete([3] * num_turbines) for name in controls})


def _state_space(attributes, controls: dict, num_turbines: int, default_bounds: dict):
    """Box per state attribute with float32 bounds: the control's own range, the default range of a measurement, or the
    (speed, direction) pair for the farm-level free-stream measurement."""
    unit = np.ones(num_turbines, dtype=np.float32)
    boxes = OrderedDict()
    for attr in attributes:
        if attr == "freewind_measurements":
            low = np.array([default_bounds["wind_speed"][0], default_bounds["wind_direction"][0]], dtype=np.float32)
            high = np.array([default_bounds["wind_speed"][1], default_bounds["wind_direction"][1]], dtype=np.float32)
        else:
            lo, hi = (controls[attr][0], controls[attr][1]) if attr in controls else default_bounds[attr]
            low, high = unit * lo, unit * hi
        boxes[attr] = spaces.Box(low, high, shape=low.shape)
    return spaces.Dict(boxes)


def clip_to_dict_space(element: dict, space) -> dict:
    """Clip every entry of ``element`` to the bounds of the Box stored under the same key."""
    for key in element:
        box = space[key]
        element[key] = np.clip(element[key], box.low, box.high)
    return element


class WindFarmMDP:
    CONTROL_SET = ["yaw", "pitch", "torque"]
    POSSIBLE_STATE_ATTRIBUTES = ["freewind_measurements", "wind_speed", "wind_direction", "yaw", "pitch", "torque"]
    DEFAULT_BOUNDS = {
        "wind_speed": [3, 28],
        "wind_direction": [0, 360],
        "yaw": [-40, 40],
        "pitch": [0, 360],
        "torque": [-1e5, 1e5],
    }
    ACTUATORS_RATE = {"yaw": 0.3, "pitch": 8}

    def __init__(self, interface: Union[BaseInterface, Type[BaseInterface]], farm_case: FarmCase, controls: dict,
                 continuous_control: bool = True, start_iter: int = 0, horizon: int = int(1e6)):
        farm_case.max_iter = horizon
        if isinstance(interface, BaseInterface):
            warn("Interface already instantiated. Simulation arguments from `Farm case` will be ignored.")
            self.interface = interface
        else:
            self.interface = interface.from_case(farm_case)
        self.num_turbines = farm_case.num_turbines
        self.continuous_control = continuous_control
        self.horizon = horizon
        self.start_iter = start_iter
        self.farm_case = farm_case

        self._check_controls(controls)
        self.controls = controls
        self.num_controls = len(controls)
        self.measures = [name for name in self.POSSIBLE_STATE_ATTRIBUTES
                         if name not in controls and name in self.interface.measure_map]
        self.state_attributes = list(controls.keys()) + self.measures

        self.action_space = _action_space(controls, self.num_turbines, continuous_control)
        self.state_space = _state_space(self.state_attributes, controls, self.num_turbines, self.DEFAULT_BOUNDS)
        self.start_state = None
        self._actuation_accumulator = self._zero_travel()

    def _zero_travel(self):
        return {name: np.zeros(self.num_turbines, dtype=np.float32) for name in self.controls}

    # -- queries --------------------------------------------------------------------------------------------------
    def get_state_powers(self):
        return self.interface.avg_powers()

    def get_accumulated_actions(self, agent=None):
        return self._actuation_accumulator.copy()  # shallow on purpose: callers see later in-place accumulation

    # -- validation -----------------------------------------------------------------------------------------------
    def _check_controls(self, control_dict: Dict):
        for name, spec in control_dict.items():
            if name not in self.CONTROL_SET:
                raise ValueError(f"Cannot control {name}. Allowed controls are {self.CONTROL_SET}")
            if name not in self.interface.CONTROL_SET:
                raise ValueError(f"Cannot control `{name}`. Interface {self.interface.__class__.__name__}"
                                 f" only allows for the following: {self.interface.CONTROL_SET}")
            if not (isinstance(spec, Iterable) and 2 <= len(spec) <= 3):
                raise TypeError(f"Wrong bounds for actuator {name}: bounds on actuators must be an iterable of the "
                                "type [lower_bound, upper_bound] or [lower_bound, upper_bound, step_size]")
            if not spec[0] < spec[1]:
                raise ValueError(f"Wrong bounds for actuator {name}: ensure that lower_bound < upper_bound")
            if len(spec) == 2:
                control_dict[name] = tuple(spec) + (1,)
                warn(f"No step size was provided for actuator {name}. Step size will default to 1.")
            elif not self.continuous_control and spec[2] <= 0:
                raise ValueError(f"Invalid step size provided for actuator {name}: it must be strictly positive")

    def _check_state(self, state: Dict):
        for attr, value in state.items():
            if attr not in self.state_attributes:
                raise ValueError(f"Unknown attribute {attr} in state dict. Accepted: {self.state_attributes}")
            if not isinstance(value, np.ndarray):
                raise TypeError(f"State attribute {attr} must be a numpy array. Received {type(value)}")
            if attr != "freewind_measurements" and value.shape != (self.num_turbines,):
                raise TypeError(f"State attribute {attr} must be of shape ({self.num_turbines},), got {value.shape}")

    # -- dynamics -------------------------------------------------------------------------------------------------
    def reset(self, seed: int = None, options: dict = None):
        rng = np.random.default_rng(seed)
        options = options or {}
        case = self.farm_case
        free = self.state_space["freewind_measurements"]
        wind_speed = wind_direction = None
        if "wind_speed" in options:
            wind_speed = options["wind_speed"]
        elif not (case.set_wind_speed or bool(case.wind_time_series is not None and len(case.wind_time_series))):
            wind_speed = np.clip(8 * rng.weibull(8), free.low[0], free.high[0])
        if "wind_direction" in options:
            wind_direction = options["wind_direction"]
        elif not (case.set_wind_direction or bool(case.wind_time_series is not None and len(case.wind_time_series))):
            wind_direction = np.clip(rng.normal(270, 20) % 360, free.low[1], free.high[1])

        self.interface.init(wind_speed, wind_direction)
        for _ in range(self.start_iter + 1):  # the warm-up solve(s) consume simulator iterations
            self.interface.update_command()
        start = OrderedDict((attr, self.interface.get_measure(attr)) for attr in self.state_attributes)
        self.start_state = clip_to_dict_space(start, self.state_space)
        self._actuation_accumulator = self._zero_travel()
        return self.start_state

    def get_controlled_state_transition(self, state: Dict, joint_action: Dict):
        if not isinstance(joint_action, dict):
            raise TypeError("Joint action must be a dictionary")
        cast = OrderedDict((k, v.astype(np.float32)) for k, v in state.items())
        state = clip_to_dict_space(cast, self.state_space)
        next_state = copy.deepcopy(state)
        for control, command in joint_action.items():
            assert control in self.controls, f"Control of `{control}` is not activated"
            command = np.array(command, np.float32)
            if self.continuous_control:
                box = self.action_space[control]
                command = np.clip(command, box.low, box.high)
            else:
                command = (command - 1) * self.controls[control][-1]
            bounds = self.state_space[control]
            next_state[control] = np.clip(state[control] + command, bounds.low, bounds.high)
            if control in self._actuation_accumulator:
                self._actuation_accumulator[control] += np.abs(command)
        return next_state

    def step_interface(self, state: Dict):
        commands = OrderedDict((control, state[control]) for control in self.controls)
        done = self.interface.update_command(**commands)
        powers = self.get_state_powers()
        for name in self.measures:
            state[name] = self.interface.get_measure(name)
        loads = self.interface.get_measure("load")
        if loads is not None:
            loads /= 1e7
        return state, powers / 1e6, loads, done

    def take_action(self, state: Dict, joint_action: Dict):
        return self.step_interface(self.get_controlled_state_transition(state, joint_action))
