"""GPU diagnostic: decode the MN-major (layout code 1) operand addressing by reading it back.
A image: word i holds i (split low/high to survive tf32's 11-bit significand); B = K-major
identity rows, so D[m][k] = the image word the tensor core fetched for A element (m, k)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_tc import run_probe, img_kmajor
NW = 16384
N = 16
Bm = torch.zeros(N, 8)
for k in range(8):
    Bm[k, k] = 1.0
b_img = img_kmajor(Bm)     # [2 chunks][16 rows][4]
bst = (N * 16, 128)
idx = torch.arange(NW)
for (lbo, sbo) in [(4096, 512), (512, 4096), (2048, 1024)]:
    words = None
    for part in range(2):
        a_img = (idx % 1024).float() if part == 0 else (idx // 1024).float()
        st = [lbo, sbo, bst[0], bst[1], 0, 0, 0, 0, 1, 0]
        D = run_probe(a_img, b_img, st, 1, N, a_mn=1, b_mn=0)
        v = D[:, :8].round().long()
        words = v if part == 0 else words + 1024 * v
    print(f"== lbo={lbo} sbo={sbo}: word offset fetched for (m, k)")
    for m in list(range(0, 12)) + [31, 32, 33, 63, 64, 96, 127]:
        print("m=%3d:" % m, " ".join("%6d" % int(words[m, k]) for k in range(8)))
