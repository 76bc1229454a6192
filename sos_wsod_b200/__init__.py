"""sos_wsod_b200 -- B200-native (sm_100a) implementation of the SoS-WSOD Stage-1 OICR+ ROI-head hot path.

Layout:
  csrc/      hand-written CUDA kernels + the C ABI (include/soswsod_b200.h) -> lib/libsoswsod_b200.so
  _lib.py    ctypes binding (fails loudly when the library is missing; no CPU fallback)
  ops.py     tensor-level wrappers over the C ABI
  layers/    wsl.layers-style operators (autograd.Function + nn.Module), e.g. RoIPool
  modeling/  the reference's plugin surface: ROIPooler, DiscriminativeAdaptionNeck, WSDDNOutputLayers,
             OICROutputLayers, OICRPlusHeads (registered in ROI_HEADS_REGISTRY)
  engine.py  the fused head step (4 views batched, manual backward) the modules delegate to
"""
__version__ = "0.1.0"
