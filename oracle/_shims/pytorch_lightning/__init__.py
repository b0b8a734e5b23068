"""Stand-in: the reference tests only use seed_everything. Test infrastructure."""
import random

import numpy as np
import torch


def seed_everything(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    return seed
