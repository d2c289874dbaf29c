"""One off-family case through resample_batch (device resident), for ncu launch lists: args in_rate out_rate quality dtype n_clips"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-soxr_b200"))
import torch, soxr
ir, orr, q, dt, n = float(sys.argv[1]), float(sys.argv[2]), sys.argv[3], sys.argv[4], int(sys.argv[5])
x = torch.randn((n, int(ir), 2), device='cuda', dtype=getattr(torch, dt))
for _ in range(3):
    y = soxr.resample_batch(x, ir, orr, q)
torch.cuda.synchronize()
print(soxr.soxr_ext.describe_path(ir, orr, soxr._quality_to_enum(q), dt))
