"""ORACLE SHIM: empty stand-in so `robust_e_nerf/external/ngp.py:34-42` imports
(the reference imports tinycudann unconditionally even for `arch: mlp`)."""
