"""morpheus_b200 -- B200-native implementation of the MorpheuS render-and-loss hot path.

Host code mirrors the reference's Python API (models.model.scene_representation, GridEncoder /
_gridencoder backend, nerfacc-shaped sampling / compositing, MorpheuS.render_rays); compute is
hand-written sm_100a CUDA behind the C ABI of include/morpheus_b200.h (libmorpheus_b200.so, built
in-tree by `python -m morpheus_b200.build`).  There is no CPU or eager fallback.
"""
from . import _lib  # noqa: F401

__all__ = ['_lib']
