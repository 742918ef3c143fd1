"""Mirror of timemachine/lib/__init__.py:12-62: the picklable integrator and barostat descriptions."""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Optional

import numpy as np

from . import custom_ops


@dataclass
class LangevinIntegrator:
    temperature: float
    dt: float
    friction: float
    masses: np.ndarray
    seed: int

    def impl(self) -> custom_ops.LangevinIntegrator:
        return custom_ops.LangevinIntegrator(self.masses, self.temperature, self.dt, self.friction, self.seed)


@dataclass
class VelocityVerletIntegrator:
    """timemachine/lib/__init__.py:24-37: cbs = -dt / masses is part of the description."""

    dt: float
    masses: np.ndarray

    cbs: np.ndarray = field(init=False)

    def __post_init__(self):
        cb = self.dt / np.asarray(self.masses, dtype=np.float64)
        cb *= -1
        self.cbs = cb

    def impl(self) -> custom_ops.VelocityVerletIntegrator:
        return custom_ops.VelocityVerletIntegrator(self.dt, self.cbs)


@dataclass
class MonteCarloBarostat:
    """timemachine/lib/__init__.py:39-62."""

    N: int
    pressure: float
    temperature: float
    group_idxs: Any
    interval: int
    seed: int
    adaptive_scaling_enabled: bool = True
    initial_volume_scale_factor: Optional[float] = None

    def impl(self, bound_potentials) -> custom_ops.MonteCarloBarostat:
        return custom_ops.MonteCarloBarostat(
            self.N,
            self.pressure,
            self.temperature,
            self.group_idxs,
            self.interval,
            bound_potentials,
            self.seed,
            self.adaptive_scaling_enabled,
            self.initial_volume_scale_factor or 0.0,  # 0.0: "use 1% of the initial box volume"
        )
