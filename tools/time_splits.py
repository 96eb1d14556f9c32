"""Developer timing: fp32 D=3 force pass at N=2^20 for the split count forced by NB_B200_SPLITS (A/B on one box)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_time import time_force
print("NB_B200_SPLITS =", os.environ.get("NB_B200_SPLITS"))
time_force(1 << 20, 3, "float32", torch.float32, reps=4)
time_force(1 << 20, 2, "float32", torch.float32, reps=3)
