"""libcontinual_b200 — B200-native (sm_100a) implementation of the per-step training hot path of RL-VIG/LibContinual.

Python here is the host-side mirror of the reference's plugin surface (`core.model.<Method>`, `core.model.backbone.<factory>`):
same constructor kwargs, method names and return tuples.  All arithmetic runs in the hand-written CUDA library
`_C/liblc_b200.so` bound through the C ABI declared in `include/lc_b200.h`; there is no CPU or eager-PyTorch fallback.
"""
__version__ = "0.1.0"
