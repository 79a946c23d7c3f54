"""Plain-struct restatement of the reference's rosparam structs for the hot path
(/root/reference/include/params/map_params.h:49-106, registration_params.h:22-30,
lidar_params.h:21-28).  Only the unit-scaling rules matter for parity."""
from dataclasses import dataclass, field

from .fixedpoint import WEIGHT_RESOLUTION


@dataclass
class MapParams:
    resolution: int = 64            # mm per voxel
    max_distance: float = 0.6       # m  -> tau = int(max_distance * 1000)
    update_distance: float = 0.5
    max_weight: int = 10            # -> * WEIGHT_RESOLUTION
    shift: float = 3.0
    size_m: tuple = (20.0, 20.0, 5.0)
    initial_weight: int = 0
    tau: int = field(init=False, default=0)
    size: tuple = field(init=False, default=(0, 0, 0))

    def __post_init__(self):
        # map_params.h:88-106 : scale_tau, scale_max_weight, scale_map_size
        self.tau = int(self.max_distance * 1000.0)
        self.max_weight = int(self.max_weight) * WEIGHT_RESOLUTION
        self.size = tuple(int(s) * 1000 // int(self.resolution) for s in self.size_m)

    @property
    def grid_size(self):
        """Side lengths the local map really gets: even sizes become s+1 (hdf5_local_map.cpp:6-8)."""
        return tuple(s if s % 2 == 1 else s + 1 for s in self.size)


@dataclass
class RegistrationParams:
    max_iterations: int = 200
    it_weight_gradient: float = 0.1
    epsilon: float = 0.03


@dataclass
class LidarParams:
    channels: int = 128
    vfov: float = 45.0
    hresolution: int = 1024


@dataclass
class Params:
    map: MapParams = field(default_factory=MapParams)
    registration: RegistrationParams = field(default_factory=RegistrationParams)
    lidar: LidarParams = field(default_factory=LidarParams)
