"""Kernel-level breakdown of the INT8 BEV backbone leg (bench.py bev_backbone_leg): torch.profiler CUDA-time table of one call."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "quantization-on-3d-object-detection_b200"))
import torch
from torch.profiler import ProfilerActivity, profile


def main():
    import qlidar
    dev = torch.device("cuda", 0)
    torch.manual_seed(6)
    m = qlidar.BaseBEVBackbone(dict(LAYER_NUMS=[5, 5], LAYER_STRIDES=[1, 2], NUM_FILTERS=[128, 256], UPSAMPLE_STRIDES=[1, 2],
                                    NUM_UPSAMPLE_FILTERS=[256, 256]), 256).to(dev).eval()
    x = torch.relu(torch.randn((4, 256, 188, 188), device=dev))
    qlidar.smoothquant(m, {}, "", 0.5, 8, 8, (torch.nn.Conv2d), qlidar.SQConv2d, [])
    with torch.no_grad():
        for _ in range(2):
            m({"spatial_features": x})
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            m({"spatial_features": x})
            torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))


if __name__ == "__main__":
    main()
