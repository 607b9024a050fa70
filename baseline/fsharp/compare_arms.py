#!/usr/bin/env python3
"""Statistical diff of the two arms' images (time_reference.sh): the reference's CPU render, the GPU render, and a second GPU
render.  The reference seeds its RNG from the clock (F9), so two renders of the same scene never agree byte for byte: the
test is that the GPU image differs from the CPU image no more than two renders of one arm differ from each other.

    compare_arms.py cpu.png gpu.png gpu2.png

Prints, per channel: mean absolute difference CPU-GPU and GPU-GPU' (gamma-corrected bytes, as the PNG holds them), the share of
bytes within 1 / 2 / 4 levels, and PASS when MAD(CPU, GPU) <= 1.5 x MAD(GPU, GPU') + 0.25 on every channel — the bound
tests/test_gpu_render.py applies between the GPU and the C++ restatement of the reference."""
import sys

import numpy as np
from PIL import Image


def load(path):
    return np.asarray(Image.open(path).convert("RGB"), dtype=np.int32)


def main():
    cpu, gpu, gpu2 = (load(p) for p in sys.argv[1:4])
    if cpu.shape != gpu.shape:
        print(f"FAIL: image sizes differ: {cpu.shape} vs {gpu.shape}")
        return 1
    mad_cg = np.abs(cpu - gpu).mean(axis=(0, 1))
    mad_gg = np.abs(gpu - gpu2).mean(axis=(0, 1))
    d = np.abs(cpu - gpu)
    print(f"image {cpu.shape[1]}x{cpu.shape[0]}")
    print(f"MAD(cpu, gpu)   per channel: {np.round(mad_cg, 3)}")
    print(f"MAD(gpu, gpu')  per channel: {np.round(mad_gg, 3)}")
    print(f"|cpu - gpu| <= 1 / 2 / 4 levels: {np.mean(d <= 1):.3f} / {np.mean(d <= 2):.3f} / {np.mean(d <= 4):.3f}; max {int(d.max())}")
    ok = bool((mad_cg <= 1.5 * mad_gg + 0.25).all())
    print("PASS" if ok else "FAIL", "- the GPU render is" + ("" if ok else " NOT") + " as close to the CPU render as to itself under another seed")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
