"""vision_mtl_b200 -- B200-native (sm_100a) implementation of the per-step multi-task hot path
of kirilllzaitsev/vision_mtl, behind the reference's own Python API.

Hot path (hand-written CUDA in ``csrc/``, C ABI in ``include/vmtl_b200.h``):
cross-stitch mixing, the MTAN attention gate, the task heads fused with their losses,
and the validation reductions.  Host code mirrors the reference modules
(``models/``, ``losses.py``, ``lit_module.py``, ``training_lit.py``).
"""
__version__ = "0.1.0"
