import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-soxr_b200"))
import torch, soxr
for (ir, orr, q) in [(48000, 8000, 'HQ'), (8000, 48000, 'HQ'), (48000, 12000, 'HQ'), (44100, 16000, 'HQ'), (48000, 44100, 'VHQ')]:
    x = torch.randn((1024, ir, 2), device='cuda')
    for _ in range(3):
        y = soxr.resample_batch(x, ir, orr, q)
    torch.cuda.synchronize()
