"""ken_burns_effect_b200 -- B200 (sm_100a) implementation of the novel-view synthesis hot path of
pierlj/ken-burns-effect behind the reference's own operator / module API.

Layout
  csrc/    hand-written CUDA kernels + the C ABI (include/kb200.h) -> lib/libkb200.so
  utils/   host-side mirror of the reference's utils/common.py, utils/pipeline.py, utils/utils.py
  models/  nn.Module mirrors (same constructors, forwards and state_dict keys as the reference)
"""
__version__ = "0.1.0"
