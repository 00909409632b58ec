#!/usr/bin/env python3
"""tools/cta_trace.py — with a -DSK_TRACE_CTAS build of the step-3 kernels (SOUNDKIT_CUDA_LIB=...): runs the bench workload
once and prints how many CTAs of the persistent fused-chain kernel were resident per SM over time (development aid)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from soundkit_b200 import capi

ctx = capi.Context(0); ctx.use_torch_stream()
S, frames = 32, int(float(sys.argv[1]) * 48000) if len(sys.argv) > 1 else 600 * 48000
n_out = ctx.batch_plan(48000, 16000, frames, capi.SEM_STREAMING)
stride = (frames * 6 + 15) // 16 * 16
raw = torch.randint(0, 256, (S, stride), dtype=torch.uint8, device="cuda")
n_pad = (n_out + 3) // 4 * 4
out = torch.empty((S, n_pad), dtype=torch.float32, device="cuda")
desc = capi.BatchDesc(raw.data_ptr(), stride, S, frames, 24, capi.FMT_SIGNED, 0, 2, capi.MODE_B, 48000, 16000, capi.SEM_STREAMING,
                      capi.MIX_SCALE_THEN_SUM, out.data_ptr(), n_pad, 0)
for _ in range(3):
    ctx.normalize_batch_dev(desc)
torch.cuda.synchronize()
lib = ctypes.CDLL(os.environ["SOUNDKIT_CUDA_LIB"])
buf = np.zeros(3 * 1024, dtype=np.uint64)
rc = lib.sk_debug_cta_trace(buf.ctypes.data_as(ctypes.c_void_p))
t = buf.reshape(1024, 3)[:444]
t0, sm, t1 = t[:, 0].astype(np.int64), t[:, 1].astype(np.int64), t[:, 2].astype(np.int64)
base = t0.min()
print("rc", rc, "kernel span us", (t1.max() - base) / 1e3)
print("CTA start offsets us: min %.1f  median %.1f  p90 %.1f  max %.1f" % tuple(np.percentile((t0 - base) / 1e3, [0, 50, 90, 100])))
print("CTA end offsets us:   min %.1f  median %.1f  p90 %.1f  max %.1f" % tuple(np.percentile((t1 - base) / 1e3, [0, 50, 90, 100])))
print("CTA durations us:     min %.1f  median %.1f  max %.1f" % tuple(np.percentile((t1 - t0) / 1e3, [0, 50, 100])))
per_sm = np.bincount(sm, minlength=148)
print("CTAs per SM histogram:", dict(zip(*np.unique(per_sm, return_counts=True))), "distinct SMs", (per_sm > 0).sum(), "max smid", sm.max())
late = (t0 - base) > 0.2 * (t1.max() - base)
print("CTAs that started after 20% of the span:", int(late.sum()))
