"""GPU diagnostic: tcgen05.mma issue->completion rate for the no-swizzle K-major layout."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_tc import run_probe, img_kmajor
g = torch.Generator().manual_seed(0)
for N in (64, 128, 256):
    K = 64
    A, B = torch.randn((128, K), generator=g) * 0.01, torch.randn((N, K), generator=g) * 0.01
    strides = [128 * 16, 128, N * 16, 128, 2 * 128 * 16, 2 * N * 16, 0, 0]
    for reps in (1, 16, 64):
        D, cyc = run_probe(img_kmajor(A), img_kmajor(B), strides, K // 8, N, reps=reps, want_cycles=True)
        n = reps * (K // 8)
        print(f"N={N:3d} mmas={n:4d} cycles={cyc:8d} -> {cyc / n:7.1f} cyc/MMA (ideal {128 * N / 256:.0f})")
