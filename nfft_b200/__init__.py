"""nfft_b200 -- B200-native (sm_100a) engine behind NFFT3's plan API.

Only what the hot path needs:
  csrc/      CUDA kernels + C ABI (libnfftcu.so) and the C host layer exporting the reference's
             own nfft_* / nfftf_* symbols (libnfft3_b200.so)
  plan.py    Python mirror of the reference plan API over libnfft3_b200.so
  cabi.py    ctypes binding of the device-pointer C ABI (include/nfftcu.h)
  dist.py    node-sharded multi-GPU driver (torch.distributed / NCCL)
There is no CPU fallback; the oracle under oracle/ is test infrastructure and is never imported here.
"""
from .plan import Plan, product_api  # noqa: F401
from . import plan_abi as flags  # noqa: F401

__all__ = ["Plan", "product_api", "flags"]
