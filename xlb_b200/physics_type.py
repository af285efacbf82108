"""Physics-type enum (reference: xlb/physics_type.py:6-8)."""

from enum import Enum, auto


class PhysicsType(Enum):
    NSE = auto()  # Navier-Stokes
    ADE = auto()  # Advection-diffusion
