#!/usr/bin/env python
"""Where the IMAGENETTE QAT step spends its time: torch.profiler kernel table for a few steps.
    python tools/qat_profile.py [--batch 64] > gpurun_out/qat_profile.txt
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learned_quantization_b200.custom_loss_terms.custom_components.custom_loss_functions import \
    sparse_categorical_crossentropy as scce  # noqa: E402
from learned_quantization_b200.training import models  # noqa: E402
from learned_quantization_b200.training.step import TrainStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    net = models.imagenette_resnet(1e-11, orientation="channelwise", kernel_storage="ohwi").cuda().train()
    x, y = models.synthetic_batch(args.batch, (224, 224, 3), "cuda")
    with torch.no_grad():
        net(x[:2])
    step = TrainStep(net, scce, optimizer="adam", lr=1e-4)
    for _ in range(5):
        step(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(x, y)
    e1.record()
    torch.cuda.synchronize()
    print(f"step {e0.elapsed_time(e1) / args.steps:.3f} ms (events)")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            step(x, y)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))


if __name__ == "__main__":
    main()
