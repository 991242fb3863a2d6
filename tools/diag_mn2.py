"""GPU diagnostic: MN-major tf32 operands in the SW128_32B layout (UMMA layout code 1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_tc import run_probe, rel

def img_mn_32b(Mt, rows_total=None, row0=0, variant=0):
    """Mt [K][MN] -> [MN/32][rows][32 words]; word (mn%4) of 16B chunk (mn/4 % 8) stored at word w ^ f(chunk)."""
    K, MN = Mt.shape
    rows_total = rows_total or K
    img = torch.zeros(MN // 32, rows_total, 32)
    idx = torch.arange(32)
    c8, w = idx // 4, idx % 4
    if variant == 0:
        dst = c8 * 4 + (w ^ (c8 % 4))          # Swizzle<2,2,2> on byte address: bits[2,4) ^= bits[4,6)
    elif variant == 1:
        dst = idx                               # no permutation
    else:
        dst = (c8 ^ 0) * 4 + w
    for a in range(MN // 32):
        blk = Mt[:, a * 32:(a + 1) * 32]
        out = torch.zeros(K, 32)
        out[:, dst] = blk
        img[a, row0:row0 + K] = out
    return img.contiguous()

g = torch.Generator().manual_seed(9)
Kt, N = 64, 128
At, Bt = torch.randn((Kt, 128), generator=g), torch.randn((Kt, N), generator=g)
ref = At.t() @ Bt
for variant in (0, 1):
    a_img, b_img = img_mn_32b(At, variant=variant), img_mn_32b(Bt, variant=variant)
    atom = Kt * 128
    for name, st in {
        "lbo=atom sbo=512 kstep=1024": [atom, 512, atom, 512, 1024, 1024, 0, 0, 1, 1],
        "lbo=512 sbo=atom kstep=1024": [512, atom, 512, atom, 1024, 1024, 0, 0, 1, 1],
        "lbo=atom sbo=1024 kstep=1024": [atom, 1024, atom, 1024, 1024, 1024, 0, 0, 1, 1],
    }.items():
        try:
            D = run_probe(a_img, b_img, st, Kt // 8, N, a_mn=1, b_mn=1)
            print("variant=%d %-30s %.3e (absmax %.3f)" % (variant, name, rel(D, ref), float(D.abs().max())))
        except AssertionError as e:
            print("variant=%d %-30s FAILED %s" % (variant, name, e))
# tap shift along K rows on the B operand (rows_total = Kt + 4)
for j in (1, 2, 3):
    Bfull = torch.randn((Kt + 4, N), generator=g)
    a_img, b_img = img_mn_32b(At), img_mn_32b(Bfull)
    atom_a, atom_b = Kt * 128, (Kt + 4) * 128
    st = [atom_a, 512, atom_b, 512, 1024, 1024, 0, j * 128, 1, 1]
    try:
        D = run_probe(a_img, b_img, st, Kt // 8, N, a_mn=1, b_mn=1)
        print("row shift j=%d: %.3e" % (j, rel(D, At.t() @ Bfull[j:j + Kt])))
    except AssertionError as e:
        print("row shift FAILED", e)
