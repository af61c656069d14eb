"""Scratch probe: per-phase time of the QR panel kernel (needs a build with EXTRA=-DLXB_QR_PROF)."""
import ctypes
import sys
import torch
import lineax_b200._native as nat
import lineax_b200._ops as ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
names = ["stage", "publish", "gsync", "read", "larfg", "pass", "gram_tail", "sync1", "lvl2+sync2", "larft"]
buf = (ctypes.c_ulonglong * 16)()
for m in (16384, 262144):
    a = torch.randn(m, n, device="cuda", dtype=torch.float32)
    ops.qr_factor(a)
    torch.cuda.synchronize()
    nat.lib.lxb_debug_qr_prof(buf)  # reset
    ops.qr_factor(a)
    torch.cuda.synchronize()
    nat.lib.lxb_debug_qr_prof(buf)
    panels = n // 32
    print(m, {k: round(buf[i] / panels / 1e3, 1) for i, k in enumerate(names)}, "us per panel")
