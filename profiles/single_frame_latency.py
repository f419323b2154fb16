import ctypes as C, numpy as np, time, sys, os
sys.path.insert(0, os.getcwd())
import spectrum_analyzer_b200 as sa
from spectrum_analyzer_b200 import cabi
lib = cabi.load(); ctx = sa.Context(0)
for n in (256, 2048, 4096, 16384):
    x = np.random.randn(n).astype(np.float32)
    fr = np.zeros(n//2+2, np.float32); va = np.zeros(n//2+2, np.float32); nb = C.c_uint32(); st = np.zeros(1, cabi.FRAME_STATS_DTYPE)
    def call():
        return lib.sa_spectrum_single(ctx._h, x.ctypes.data, n, 44100, 0, 0.0, 0.0, 1, 2, fr.ctypes.data, va.ctypes.data, C.byref(nb), st.ctypes.data)
    for _ in range(50): assert call() == 0
    t0 = time.perf_counter()
    for _ in range(2000): call()
    dt = (time.perf_counter() - t0) / 2000
    print(n, "single-frame call latency %.1f us" % (dt * 1e6))
