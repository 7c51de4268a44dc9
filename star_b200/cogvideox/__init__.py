"""CogVideoX-5B DiT layer (STAR's patched block) on the sm_100a kernels."""
from .dit_block import DiTLayer  # noqa: F401
