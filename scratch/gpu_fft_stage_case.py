import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-soxr_b200"))
import torch, soxr
ir, orr, q = (int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]) if len(sys.argv) > 3 else (44100, 32000, 'HQ')
x = torch.randn((512, ir, 2), device='cuda')
for _ in range(6):
    y = soxr.resample_batch(x, ir, orr, q)
torch.cuda.synchronize()
print(y.shape)
