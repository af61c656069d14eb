"""Accuracy of the large-QR factors against a float64 LAPACK factorisation (run once with the
default tensor-core update and once with LXB_QR_TC=0)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lineax_b200._ops as ops

m, n = 32768, 1024
rng = np.random.default_rng(0)
a = (rng.standard_normal((m, n)) / np.sqrt(m)).astype(np.float32)
b = rng.standard_normal(m).astype(np.float32)
aq, taus = ops.qr_factor(torch.as_tensor(a).cuda()[None])
x = ops.qr_solve(aq, taus, torch.as_tensor(b).cuda()[None], False)[0].cpu().numpy()
r = np.triu(aq[0, :n].cpu().numpy().astype(np.float64))
r64 = np.linalg.qr(a.astype(np.float64), mode="r")
s = np.sign(np.diag(r)) * np.sign(np.diag(r64))
r64 = r64 * s[:, None]
x64 = np.linalg.lstsq(a.astype(np.float64), b.astype(np.float64), rcond=None)[0]
print("LXB_QR_TC=%s  max|R - R64| / max|R64| = %.3e   ||R - R64||_F / ||R64||_F = %.3e   max|x - x64| / max|x64| = %.3e"
      % (os.environ.get("LXB_QR_TC", "default(1)"), np.abs(r - r64).max() / np.abs(r64).max(),
         np.linalg.norm(r - r64) / np.linalg.norm(r64), np.abs(x - x64).max() / np.abs(x64).max()))
