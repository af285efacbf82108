"""Force operators: momentum exchange on a no-slip boundary, exact-difference body force (reference: xlb/operator/force/)."""

from xlb_b200._exports import export

export(globals(), __name__, {"momentum_transfer": ["MomentumTransfer"], "exact_difference_force": ["ExactDifference"]})
