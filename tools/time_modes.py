"""Developer timing helper: force-kernel rate per precision mode (D=2 disk and D=3 box), one line per case."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.quick_time import time_force  # noqa: E402

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
    modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["float32", "float16", "bfloat16", "int8_sim", "int4_sim", "custom"]
    for dim in (2, 3):
        for mode in modes:
            time_force(n, dim, mode, torch.float32)
