"""Collision operators in scope: BGK and KBC (ForcedCollision / SmagorinskyLESBGK are not)."""

from xlb_b200._exports import export

export(globals(), __name__, {"collision": ["Collision"], "bgk": ["BGK"], "kbc": ["KBC"]})
