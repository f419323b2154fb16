"""cfg4 (BASELINE.json configs[3]) for profiling: 16384-point BH4, Range(50,15000), zero_to_one, min/max stats."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch, spectrum_analyzer_b200 as sa
from spectrum_analyzer_b200 import cabi
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n = 16384
lib = cabi.load(); dev = torch.device("cuda", 0); stream = torch.cuda.current_stream(dev)
ctx = sa.Context(0, stream.cuda_stream); chk = sa.api._check
x = torch.empty(frames * n, dtype=torch.float32, device=dev); chk(lib.sa_generate_noise(ctx._h, x.data_ptr(), 0, x.numel(), 1))
lim = sa.FrequencyLimit.Range(50.0, 15000.0)
blo, nb = lim.to_bins(n, 44100)
vals = torch.empty(frames * nb, dtype=torch.float32, device=dev); stats = torch.empty(frames * 32, dtype=torch.uint8, device=dev)
for _ in range(4):
    chk(lib.sa_spectrum_batched(ctx._h, x.data_ptr(), frames, n, n, 44100, lim.kind, lim.lo, lim.hi, cabi.WINDOW_BLACKMAN_HARRIS_4TERM,
                                cabi.SCALING_ZERO_TO_ONE, 3, vals.data_ptr(), nb, stats.data_ptr(), None, None))
torch.cuda.synchronize(); print("ok")
