"""gym_auv_b200 -- B200-native batched simulator for gym-auv's per-step hot path.

Only what the path needs lives here: ``csrc/`` (sm_100a kernels + C ABI), the ctypes
binding, the host-side mirror of the reference's env interface (``Config``, scenario
plug-ins, ``AUVVecEnv``, the single-env ``gym.Env``-shaped facade) and the path-bank
builder.  See DESIGN.md / INTEGRATION.md.
"""
from .config import Config, EpisodeConfig, SimulationConfig, VesselConfig, RenderingConfig, lidar_config, effective_reference_config  # noqa: F401
from . import scenarios  # noqa: F401
from .scenarios import SCENARIOS, ScenarioSet  # noqa: F401

__all__ = [
    "Config", "EpisodeConfig", "SimulationConfig", "VesselConfig", "RenderingConfig", "lidar_config",
    "effective_reference_config", "scenarios", "SCENARIOS", "ScenarioSet", "AUVVecEnv", "AUVEnv", "make",
]


def __getattr__(name):  # torch / CUDA only needed once the compute path is touched
    if name == "AUVVecEnv":
        from .vec_env import AUVVecEnv

        return AUVVecEnv
    if name in ("AUVEnv", "make"):
        from . import env as _env

        return getattr(_env, name)
    raise AttributeError(name)
