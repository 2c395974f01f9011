"""Hand-written sm_100a kernels and their Python launchers.

``csrc/cuda/*.cu`` -> ``blades_b200/_cuda.so`` (extern "C" launchers, loaded with ctypes;
kernels take raw device pointers + the current CUDA stream), ``csrc/host/*.cpp`` ->
``blades_b200/_host.so`` (CPU-side selectors / data assembly).  Build with
``python -m blades_b200.ops.build`` (or ``__graft_entry__.build()``).
On a GPU box a missing ``_cuda.so`` is a hard error (no silent eager fallback)."""
