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
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--hop", type=int, default=0, help="samples between frames (default: n)")
ap.add_argument("--scaling", type=int, default=2, help="sa_scaling_kind (2 = divide_by_N_sqrt)")
ap.add_argument("--window", type=int, default=1, help="sa_window_kind (1 = Hann)")
ap.add_argument("--limit", nargs=3, default=["0", "0", "0"], help="limit kind, lo, hi")
ap.add_argument("names", nargs="+")
a = ap.parse_args()
root = os.getcwd()
n, frames = a.n, a.frames
hop = a.hop or n
lk, llo, lhi = int(a.limit[0]), float(a.limit[1]), float(a.limit[2])
x = torch.empty((frames - 1) * hop + n, dtype=torch.float32, device="cuda")
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
        r = lib.sa_spectrum_batched(ctx, x.data_ptr(), frames, n, hop, 44100, lk, llo, lhi, a.window, a.scaling, a.stats,
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
    bpf = 4 * hop + 4 * nb.value + 32          # algorithmic bytes per frame
    print(f"{name:16s} {fps / 1e6:8.2f} M spectra/s  frac {fps * bpf / 1e9 / peak:.3f}  (n={n} hop={hop} scaling={a.scaling} "
          f"window={a.window} limit={lk} stats={a.stats} bins={nb.value})", flush=True)
