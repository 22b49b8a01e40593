"""Synthetic keyframe descriptor sets for the loop-closure tests (the generator lives with the other synthetic inputs)."""
from svin_b200.synthetic_loop import make_keyframes  # noqa: F401
