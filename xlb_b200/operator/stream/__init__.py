"""Streaming operator namespace."""

from xlb_b200._exports import export

export(globals(), __name__, {"stream": ["Stream"]})
