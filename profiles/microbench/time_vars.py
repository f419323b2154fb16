"""Time experiment builds of the library in ONE process (the headline workload, device-resident).
usage: python profiles/microbench/time_vars.py [--frames F] [--stats M] [--steps K] name1 name2 ...
       name = variant in profiles/microbench/variants/libsa_<name>.so, or "main" for the shipped library.
Timing: CUDA events on the context's own stream (sa_ctx_sync brackets), 3 warm-ups, K launches, inputs >> L2."""
import argparse, ctypes as C, os, sys
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=393216)
ap.add_argument("--stats", type=int, default=7)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("names", nargs="+")
a = ap.parse_args()
root = os.getcwd()
n, frames = 4096, a.frames
x = torch.empty(frames * n, dtype=torch.float32, device="cuda")
out = torch.empty(frames * (n // 2 + 1), dtype=torch.float32, device="cuda")
st = torch.empty(frames * 8, dtype=torch.int32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
peak = 6542.1
try:
    import json
    peak = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
first = True
for name in a.names:
    path = (os.path.join(root, "spectrum_analyzer_b200/lib/libspectrum_analyzer_cuda.so") if name == "main"
            else os.path.join(root, f"profiles/microbench/variants/libsa_{name}.so"))
    if not os.path.exists(path):
        print(name, "not built"); continue
    lib = C.CDLL(path)
    ctx = C.c_void_p()
    assert lib.sa_ctx_create_on_stream(0, C.c_void_p(stream), C.byref(ctx)) == 0
    lib.sa_generate_noise.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64]
    if first:
        assert lib.sa_generate_noise(ctx, x.data_ptr(), 0, x.numel(), 0x5A17) == 0
        first = False
    lib.sa_spectrum_batched.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_int,
                                        C.c_float, C.c_float, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_uint64,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
    blo, nb = C.c_uint32(), C.c_uint32()
    def run():
        r = lib.sa_spectrum_batched(ctx, x.data_ptr(), frames, n, n, 44100, 0, 0.0, 0.0, 1, 2, a.stats,
                                    out.data_ptr(), n // 2 + 1, st.data_ptr() if a.stats else None, C.byref(blo), C.byref(nb))
        assert r == 0, r
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps): run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    fps = frames / ms * 1e3
    print(f"{name:16s} {fps / 1e6:8.2f} M spectra/s  frac {fps * 24612 / 1e9 / peak:.3f}  (stats={a.stats})", flush=True)
