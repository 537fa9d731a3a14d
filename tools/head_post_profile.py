"""ncu driver for the CenterHead post-processing kernels: python tools/head_post_profile.py (bench.head_post_leg, 5 calls)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "quantization-on-3d-object-detection_b200"))
import torch

import bench

if __name__ == "__main__":
    torch.cuda.set_device(0)
    print(bench.head_post_leg(torch.device("cuda", 0), iters=5))
