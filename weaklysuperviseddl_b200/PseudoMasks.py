"""Alias: the reference imports `PseudoMasks` (AlternatingDirectionBoundaryLoss.py:9, Abalations.py:3) although
its file is spelt PsuedoMasks.py."""
from .PsuedoMasks import *  # noqa: F401,F403
from .PsuedoMasks import delete_dir_recursive, generate_pseudo_masks, keep_largest  # noqa: F401
