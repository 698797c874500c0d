"""surface_sampling_b200 — B200-native energy/force + relaxation engine for the VSSR-MC hot path.

Only what the path needs lives here: ``csrc/`` (hand-written sm_100a CUDA + the C ABI),
``_lib`` (ctypes binding), ``engine`` (batched device-resident engines) and the host-side mirrors
of the reference's calculator / relax interface (``calculators``, ``dynamics``, ``mc``).
"""
__version__ = "0.1.0"
