"""One configuration, a few launches (for ncu): python profiles/run_one_config.py N HOP FRAMES SCALING WINDOW STATS [LIMIT_KIND LO HI]"""
import os
import sys

sys.path.insert(0, os.getcwd())
import torch

import spectrum_analyzer_b200 as sa
from spectrum_analyzer_b200 import cabi

n, hop, frames, sc, win, st = (int(a) for a in sys.argv[1:7])
lk, lo, hi = (int(sys.argv[7]), float(sys.argv[8]), float(sys.argv[9])) if len(sys.argv) > 9 else (0, 0.0, 0.0)
lib = cabi.load()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)
ctx = sa.Context(0, stream.cuda_stream)
chk = sa.api._check
x = torch.empty((frames - 1) * hop + n, dtype=torch.float32, device=dev)
chk(lib.sa_generate_noise(ctx._h, x.data_ptr(), 0, x.numel(), 1))
nb = sa.FrequencyLimit(lk, lo, hi).to_bins(n, 44100)[1]
vals = torch.empty(frames * nb, dtype=torch.float32, device=dev)
stats = torch.empty(frames * 32, dtype=torch.uint8, device=dev)
for _ in range(4):
    chk(lib.sa_spectrum_batched(ctx._h, x.data_ptr(), frames, n, hop, 44100, lk, lo, hi, win, sc, st, vals.data_ptr(), nb,
                                stats.data_ptr(), None, None))
torch.cuda.synchronize()
print("ok", ctx.last_kernel)
