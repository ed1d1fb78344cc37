"""vierkant_b200 -- B200-native (sm_100a) BCn texture block encoders behind vierkant::bcn::compress()'s contract.

  csrc/      CUDA kernels + the C ABI (libvierkant_bcn_cuda, include/vierkant_bcn_cuda.h)
  capi.py    ctypes binding of that ABI
  compress.py  Python mirror of vierkant::bcn::compress / compress_info_t / compress_result_t (for tests and bench)
  synth.py   deterministic synthetic textures
"""
__all__ = ["capi", "synth", "build"]
