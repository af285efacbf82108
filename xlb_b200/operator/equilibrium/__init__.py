"""Equilibrium operators."""

from xlb_b200._exports import export

export(globals(), __name__, {"equilibrium": ["Equilibrium"], "quadratic_equilibrium": ["QuadraticEquilibrium"]})
