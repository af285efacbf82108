"""Moments of the populations: density, velocity, momentum flux."""

from xlb_b200._exports import export

export(globals(), __name__, {"macroscopic": ["Macroscopic"], "second_moment": ["SecondMoment"], "zero_moment": ["ZeroMoment"], "first_moment": ["FirstMoment"]})
