"""myconvnet_b200 — B200-native backend for the data-parallel training step of MyConvNet models.

Layout
  csrc/ + libmcn.so   hand-written sm_100a CUDA kernels behind a C ABI (include/mcn.h)
  lib.py              ctypes binding (no fallback)
  tfshim/, convnet.py host-side mirror of the reference's model-definition API
  graph.py, plan.py   static layer graph -> fused launch sequence + memory plan (CPU-only logic)
  engine.py           executes a plan: training step, optimiser, multi-GPU collectives
  loader.py           imports the reference's model files unchanged against the facade
"""
__version__ = "0.1.0"
