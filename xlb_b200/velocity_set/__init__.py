"""Lattices: base class and the three velocity sets of the reference (D2Q9, D3Q19, D3Q27)."""

from xlb_b200._exports import export

export(globals(), __name__, {"velocity_set": ["VelocitySet"], "d2q9": ["D2Q9"], "d3q19": ["D3Q19"], "d3q27": ["D3Q27"]})
