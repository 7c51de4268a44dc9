"""setup_seed (reference: video_to_video/utils/seed.py:9-14)."""
import random

import numpy as np
import torch


def setup_seed(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
