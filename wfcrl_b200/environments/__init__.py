"""``envs.make`` / ``envs.make_vec`` / ``envs.list_envs`` and the case classes, re-exported under the names the
reference's ``wfcrl.environments`` package uses."""
from . import data_cases as _cases
from . import registration as _registry

FarmCase = _cases.FarmCase
FlorisCase = _cases.FlorisCase
make = _registry.make
make_vec = _registry.make_vec
list_envs = _registry.list_envs

__all__ = ["FarmCase", "FlorisCase", "make", "make_vec", "list_envs"]
