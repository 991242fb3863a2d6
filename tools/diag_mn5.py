"""GPU diagnostic: MN-major (layout 1) addressing when the descriptor start is shifted by whole 128 B rows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_tc import run_probe, img_kmajor
NW, N = 16384, 16
Bm = torch.zeros(N, 8)
for k in range(8):
    Bm[k, k] = 1.0
b_img = img_kmajor(Bm)
idx = torch.arange(NW)
lbo, sbo = 8192, 512
for a_off in (0, 128, 256, 384, 512, 640):
    words = None
    for part in range(2):
        a_img = (idx % 1024).float() if part == 0 else (idx // 1024).float()
        st = [lbo, sbo, N * 16, 128, 0, 0, a_off, 0, 1, 0]
        D = run_probe(a_img, b_img, st, 1, N, a_mn=1, b_mn=0)
        v = D[:, :8].round().long()
        words = v if part == 0 else words + 1024 * v
    print(f"a_off={a_off:4d}B (row {a_off // 128}):", " | ".join("m=%d: %s" % (m, " ".join("%5d" % int(words[m, k]) for k in range(8))) for m in (0, 8)))
