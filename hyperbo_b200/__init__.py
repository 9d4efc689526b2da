"""hyperbo_b200 -- B200-native GP pre-training / inference engine behind the
hyperbo.gp_utils GP / kernel / objectives API.

Package map (mirrors the reference's module names for the hot path only):
  basics.{definitions, linalg, params_utils, data_utils}
  gp_utils.{kernel, mean, objectives, gp, utils}
  bo_utils.{acfun, const}
  engine      -- host driver of the C-ABI library (libhyperbo_b200.so)
  csrc/       -- sm_100a CUDA kernels + the C ABI + the pybind11 forwarding layer
"""
__version__ = "0.1.0"
