"""Mirror of timemachine/lib/__init__.py:12-21 for the hot path: the picklable integrator description."""

from __future__ import annotations

from dataclasses import dataclass

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
