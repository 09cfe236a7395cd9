"""consistent_depth_b200 — B200-native (sm_100a) test-time depth fine-tuning hot path.

Host-side mirror of facebookresearch/consistent_depth's plugin surface for the
fine-tuning path (monodepth/depth_model{,_registry}.py, loss/*, optimizer/,
utils/geometry.py, depth_fine_tuning.py) over hand-written CUDA in
libcvd_sm100.so (C-ABI in include/cvd.h).  There is no CPU fallback: using a
kernel-backed op without the built library / a CUDA device raises.
"""
__version__ = "0.1.0"
