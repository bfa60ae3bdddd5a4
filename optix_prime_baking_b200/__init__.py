"""optix_prime_baking_b200 — B200-native ambient-occlusion baker (drop-in for the
optix_prime_baking bake path).  The product is libaobake.so (CUDA, sm_100a) behind the
C-ABI of include/aobake.h; this package is its Python host-side mirror plus scene helpers."""
from . import scenes  # noqa: F401

__all__ = ["scenes", "api", "build"]
