"""Physics-type enum (reference: xlb/physics_type.py:6-8): NSE = Navier-Stokes, ADE = advection-diffusion."""

from enum import Enum

PhysicsType = Enum("PhysicsType", ["NSE", "ADE"])
