"""Force evaluation: momentum exchange on a no-slip boundary."""

from xlb_b200._exports import export

export(globals(), __name__, {"momentum_transfer": ["MomentumTransfer"]})
