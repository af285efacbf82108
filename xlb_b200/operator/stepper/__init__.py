"""Steppers: the base class and the fused incompressible Navier-Stokes stepper."""

from xlb_b200._exports import export

export(globals(), __name__, {"stepper": ["Stepper"], "nse_stepper": ["IncompressibleNavierStokesStepper"]})
