"""Configuration tree with the reference's field names (gym_auv/config.py:13-119).

Differences from the reference, all deliberate:
  * sub-configs are created with ``default_factory`` so two ``Config()`` objects never
    share state (the reference's class-level defaults make ``DEBUG_CONFIG``'s mutation
    leak into every config, ``gym_auv/__init__.py:24-26``);
  * ``effective_reference_config()`` reproduces that leak explicitly for users who want
    the values the reference actually runs with after ``import gym_auv``
    (``t_step_size = 0.5``, ``min_goal_distance = 0.1``).
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import Tuple, Union


@dataclass
class EpisodeConfig:
    min_cumulative_reward: float = -2000.0  # config.py:15-17
    max_timesteps: int = 10000  # config.py:19
    min_goal_distance: float = 5.0  # config.py:20-22
    min_path_progress: float = 0.99  # config.py:23


@dataclass
class SimulationConfig:
    t_step_size: float = 1.0  # config.py:28
    sensor_frequency: float = 1.0  # config.py:29-31 (unused on the path at HEAD)
    observe_frequency: float = 1.0  # config.py:32-34 (unused on the path at HEAD)


@dataclass
class VesselConfig:
    thrust_max_auv: float = 2.0
    moment_max_auv: float = 0.15
    vessel_width: float = 1.255
    feasibility_width_multiplier: float = 5.0
    look_ahead_distance: int = 300
    render_distance: Union[int, str] = 300
    include_original_observations: bool = False
    use_relative_vectors: bool = True
    use_lidar: bool = False  # config.py:52-55 -- the reference default is OFF
    sensor_interval_load_obstacles: int = 25
    n_sensors_per_sector: int = 20
    n_sectors: int = 9
    sensor_use_feasibility_pooling: bool = False
    sensor_use_velocity_observations: bool = False
    sensor_rotation: bool = False
    sensor_range: float = 150.0
    sensor_log_transform: bool = True
    use_dict_observation: bool = False

    @property
    def n_sensors(self) -> int:  # config.py:75-78
        return self.n_sensors_per_sector * self.n_sectors

    @property
    def lidar_shape(self) -> Tuple[int, int]:  # config.py:80-87
        return (3 if self.sensor_use_velocity_observations else 1, self.n_sensors)

    @property
    def n_lidar_observations(self) -> int:  # config.py:89-91
        return self.lidar_shape[0] * self.lidar_shape[1]

    @property
    def dense_observation_size(self) -> int:  # config.py:93-98
        return 6


@dataclass
class RenderingConfig:
    show_indicators: bool = True
    autocamera3d: bool = True


@dataclass
class Config:
    episode: EpisodeConfig = field(default_factory=EpisodeConfig)
    simulation: SimulationConfig = field(default_factory=SimulationConfig)
    vessel: VesselConfig = field(default_factory=VesselConfig)
    rendering: RenderingConfig = field(default_factory=RenderingConfig)

    def __iter__(self):  # config.py:118-119
        return iter(dataclasses.fields(self))

    def copy(self) -> "Config":
        return dataclasses.replace(
            self,
            episode=dataclasses.replace(self.episode),
            simulation=dataclasses.replace(self.simulation),
            vessel=dataclasses.replace(self.vessel),
            rendering=dataclasses.replace(self.rendering),
        )


def lidar_config(**vessel_overrides) -> Config:
    """Declared defaults with the 180-ray LiDAR switched on (the BASELINE workload)."""
    cfg = Config()
    cfg.vessel.use_lidar = True
    for k, v in vessel_overrides.items():
        setattr(cfg.vessel, k, v)
    return cfg


def effective_reference_config() -> Config:
    """What every reference config looks like after ``import gym_auv`` on Python <= 3.10
    (shared sub-config instances + DEBUG_CONFIG mutation, gym_auv/__init__.py:24-26)."""
    cfg = Config()
    cfg.simulation.t_step_size = 0.5
    cfg.episode.min_goal_distance = 0.1
    return cfg
