"""elm_physics_b200 — B200-native drop-in for the hot path behind `Physics.simulate` of w0rm/elm-physics.

Layout: `csrc/` CUDA kernels + the C ABI (include/elm_physics_b200.h); `engine.py` ctypes binding of the
built library; `physics.py`/`shapes.py`/`math3d.py` host-side mirror of the reference's construction-time
API; `pack.py`/`abi.py` pure marshalling; `scenes.py` the BASELINE.json configs as synthetic inputs.
"""
