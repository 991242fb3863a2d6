"""GPU diagnostic: brute-force the MN-major tf32 (layout code 1, SW128_32B) data arrangement."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_tc import run_probe, rel

def build(Mt, arrangement, swz):
    """Mt [K][MN]. Returns image tensor (flat float32) and dict of natural strides."""
    K, MN = Mt.shape
    na, ka = MN // 32, K // 4
    kk = torch.arange(K)[:, None].expand(K, MN)
    mn = torch.arange(MN)[None, :].expand(K, MN)
    a, m32 = mn // 32, mn % 32
    c8, w = m32 // 4, m32 % 4
    if swz == 0:   w2, c2 = w ^ (c8 % 4), c8
    elif swz == 1: w2, c2 = w, c8
    elif swz == 2: w2, c2 = w ^ (kk % 4), c8
    elif swz == 3: w2, c2 = w, c8 ^ (kk % 4)
    elif swz == 4: w2, c2 = w, c8 ^ (kk % 8)
    elif swz == 5: w2, c2 = w ^ (c8 // 2), c8
    within = c2 * 4 + w2
    if arrangement == "A":      # [mn_atom][k][32]
        off = a * (K * 32) + kk * 32 + within
    else:                        # "B": [k_atom][mn_atom][4][32]
        off = (kk // 4) * (na * 128) + a * 128 + (kk % 4) * 32 + within
    img = torch.zeros(16384)   # 64 KB: every candidate stride combination stays inside the image
    img[off.flatten()] = Mt.flatten()
    return img

g = torch.Generator().manual_seed(9)
Kt, N = 16, 128
At, Bt = torch.randn((Kt, 128), generator=g), torch.randn((Kt, N), generator=g)
ref = At.t() @ Bt
results = []
for arr in ("A", "B"):
    for swz in range(6):
        a_img, b_img = build(At, arr, swz), build(Bt, arr, swz)
        cand = sorted(set([128, 512, 1024, 2048, Kt * 128, 4 * 512]))
        for lbo, sbo in itertools.product(cand, cand):
            for kstep in sorted(set([1024, 2 * sbo, 2 * lbo])):
                st = [lbo, sbo, lbo, sbo, kstep, kstep, 0, 0, 1, 1]
                try:
                    D = run_probe(a_img, b_img, st, Kt // 8, N, a_mn=1, b_mn=1)
                    results.append((rel(D, ref), arr, swz, lbo, sbo, kstep))
                except AssertionError:
                    pass
results.sort()
for r in results[:12]:
    print("err %.3e arr=%s swz=%d lbo=%d sbo=%d kstep=%d" % r)
print("tried", len(results))
