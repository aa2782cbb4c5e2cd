"""Refine for the checkpoints released with the original 3D Ken Burns paper -- mirror of
models/disparity_refinement_pretrained.py:80-128: same topology, Basic blocks WITH residual / 1x1 shortcut."""
from .disparity_refinement import Refine as _Refine


class Refine(_Refine):
    SHORTCUT = 'auto'
