"""One basis_at launch for ncu.  usage: python profiles/run_basis_one.py values|grad [n]"""
import sys
sys.path.insert(0, ".")
import torch
import __graft_entry__ as ge
sk = ge.load_package()
mode = sys.argv[1] if len(sys.argv) > 1 else "values"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
b4 = sk.smolyak_basis(sk.Chebyshev, sk.InteriorGrid2(), sk.SmolyakParameters(4), 6)
u = torch.rand((n, 6), device="cuda", dtype=torch.float64) * 2 - 1
for _ in range(2):
    out = sk.basis_at(b4, sk.gradient(6)(u) if mode == "grad" else u)
torch.cuda.synchronize()
print(out.shape)
