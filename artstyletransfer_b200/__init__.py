"""artstyletransfer_b200 — the pyramid Gatys-loss hot path of irenemizus/ArtStyleTransfer as hand-written
sm_100a CUDA (libast_sm100.so, C ABI in include/ast_sm100.h) behind the reference's own Python call surface.

Modules mirror the reference's names: math_utils, neural_nets, neural_style_transfer, config.
There is no CPU fallback: ops raise if the shared library is missing or a tensor is not on a CUDA device.
"""
__version__ = '0.1.0'
