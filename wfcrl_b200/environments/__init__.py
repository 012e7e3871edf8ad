from .data_cases import FarmCase, FlorisCase  # noqa: F401
from .registration import list_envs, make, make_vec  # noqa: F401
