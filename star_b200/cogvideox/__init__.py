"""STAR's CogVideoX-5B DiT (cogvideox-based/) on the sm_100a kernels: one layer (DiTLayer) and the whole
DiffusionTransformer (patch embed, 42 layers with LoRA merged, final layer)."""
from .dit_block import DiTLayer  # noqa: F401
from .model import DiffusionTransformer, dit_manifest, rope_tables  # noqa: F401
