"""Diagnostic: a few QR calls for an ncu launch list (not part of the product)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from renormalizer_b200 import ops, _lib
from renormalizer_b200.backend import asxp
_lib.get()
rng = np.random.default_rng(0)
m, n = int(sys.argv[1]), int(sys.argv[2])
a = asxp(rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n)))
for _ in range(3):
    q, r = ops.qr(a)
torch.cuda.synchronize()
l, q2 = ops.qr(asxp(rng.standard_normal((n, m)) + 1j * rng.standard_normal((n, m))), lq=True)
torch.cuda.synchronize()
