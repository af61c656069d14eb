"""Accuracy of the large-QR factors against a float64 LAPACK factorisation (run under the env switches
LXB_QR_TWOLEVEL / LXB_QR_TC to compare paths); prints the error of R per 128-row block."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lineax_b200._ops as ops

m, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32768, 1024)
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
tag = " ".join(f"{k}={os.environ[k]}" for k in ("LXB_QR_TWOLEVEL", "LXB_QR_TC") if k in os.environ) or "default"
print("%s %dx%d  max|R - R64| / max|R64| = %.3e   ||R - R64||_F / ||R64||_F = %.3e   max|x - x64| / max|x64| = %.3e"
      % (tag, m, n, np.abs(r - r64).max() / np.abs(r64).max(),
         np.linalg.norm(r - r64) / np.linalg.norm(r64), np.abs(x - x64).max() / np.abs(x64).max()))
err = np.abs(r - r64) / np.abs(r64).max()
print("  per 128-row block of R:", " ".join("%.1e" % err[i:i + 128].max() for i in range(0, n, 128)))
