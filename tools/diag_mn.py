"""GPU diagnostic: MN-major tf32 operands with the 128B swizzle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_tc import run_probe, rel

def img_mn_sw128(Mt, swz=True):
    """Mt: [K][MN] (k rows, MN contiguous) -> image [MN/32][K][32] with 16B-chunk XOR swizzle by (k%8)."""
    K, MN = Mt.shape
    img = torch.zeros(MN // 32, K, 32)
    for a in range(MN // 32):
        blk = Mt[:, a * 32:(a + 1) * 32].reshape(K, 8, 4)       # [k][chunk][4]
        if swz:
            out = torch.zeros_like(blk)
            for k in range(K):
                for c in range(8):
                    out[k, c ^ (k % 8)] = blk[k, c]
            blk = out
        img[a] = blk.reshape(K, 32)
    return img.contiguous()

g = torch.Generator().manual_seed(9)
Kt, N = 64, 128
At, Bt = torch.randn((Kt, 128), generator=g), torch.randn((Kt, N), generator=g)
ref = At.t() @ Bt
for swz in (True, False):
    a_img, b_img = img_mn_sw128(At, swz), img_mn_sw128(Bt, swz)
    atom_mn = Kt * 128          # bytes between 32-wide MN atoms (all k rows of one atom first)
    for name, st in {
        "SW128 lbo=atom_mn sbo=1024": [atom_mn, 1024, atom_mn, 1024, 1024, 1024, 0, 0, 2, 2],
        "SW128 lbo=1024 sbo=atom_mn": [1024, atom_mn, 1024, atom_mn, 1024, 1024, 0, 0, 2, 2],
        "NONE-code lbo=atom_mn sbo=1024": [atom_mn, 1024, atom_mn, 1024, 1024, 1024, 0, 0, 0, 0],
    }.items():
        try:
            D = run_probe(a_img, b_img, st, Kt // 8, N, a_mn=1, b_mn=1)
            print("swz=%d %-32s %.3e (absmax %.3f)" % (swz, name, rel(D, ref), float(D.abs().max())))
        except AssertionError as e:
            print("swz=%d %-32s FAILED %s" % (swz, name, e))
# sanity: K-major SW128 for both (image [rows][32 k] per 8-row atom, chunk ^ row%8), K=32 per 128B row
def img_k_sw128(Mat):
    R, K = Mat.shape  # K multiple of 32
    img = torch.zeros(K // 32, R, 32)
    for kb in range(K // 32):
        blk = Mat[:, kb * 32:(kb + 1) * 32].reshape(R, 8, 4)
        out = torch.zeros_like(blk)
        for r in range(R):
            for c in range(8):
                out[r, c ^ (r % 8)] = blk[r, c]
        img[kb] = out.reshape(R, 32)
    return img.contiguous()
A, B = At.t().contiguous(), Bt.t().contiguous()   # [128][Kt], [N][Kt]
st = [16, 1024, 16, 1024, 32, 32, 0, 0, 2, 2]      # K-major SW128: SBO = 8 rows * 128 B; K-step = +32 B inside the row
try:
    D = run_probe(img_k_sw128(A[:, :32].contiguous()), img_k_sw128(B[:, :32].contiguous()), st, 4, N)
    print("K-major SW128 (K=32): %.3e" % rel(D, A[:, :32] @ B[:, :32].t()))
except AssertionError as e:
    print("K-major SW128 FAILED", e)
