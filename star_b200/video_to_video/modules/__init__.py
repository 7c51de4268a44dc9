"""Operator surface of the STAR denoiser (mirrors video_to_video/modules/__init__.py:1-2, which
star-exports embedder.py and unet_v2v.py, including the imported names callers rely on)."""
from .embedder import *      # noqa: F401,F403
from .unet_v2v import *      # noqa: F401,F403
