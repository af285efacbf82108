"""Collision operators: BGK, KBC, SmagorinskyLESBGK and the ForcedCollision wrapper (reference: xlb/operator/collision/)."""

from xlb_b200._exports import export

export(
    globals(),
    __name__,
    {"collision": ["Collision"], "bgk": ["BGK"], "kbc": ["KBC"], "smagorinsky_les_bgk": ["SmagorinskyLESBGK"], "forced_collision": ["ForcedCollision"]},
)
