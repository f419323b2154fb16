"""Regenerates tests/golden/*.json.

Two kinds of fixtures:
  reference_known_answers.json  the known answers the reference's OWN tests hold for the hot path, transcribed by hand from
                                /root/reference (file:line next to each entry).  The reference is Rust (microfft, libm crates;
                                no cargo/rustc in this image), so it cannot be run here: these constants are all it offers.
  oracle_vectors.json           small seeded input/output vectors of the CPU oracle (oracle/sa_oracle.c), frozen so that a
                                change of the oracle itself is noticed: the oracle is pinned to the reference's known answers
                                above and to an independent f64 DFT (tests/test_oracle_golden.py), these vectors pin it to
                                itself over time.  usage: python tests/golden/make_golden.py   (needs gcc, no GPU)
"""
import json, os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

KNOWN = {
    "windows": [
        {"ref": "src/windows.rs:156-164", "kind": "hamming", "input": [1.0, 1.0, 1.0, 1.0],
         "expect": [0.08, 0.77, 0.77, 0.08], "atol": 1e-5},
        {"ref": "src/windows.rs:166-174", "kind": "blackman_harris_4term", "input": [2.0, 2.0, 2.0, 2.0],
         "expect": [0.00012, 1.04115, 1.04115, 0.00012], "atol": 1e-5},
    ],
    "scale_to_zero_to_one": {"ref": "src/scaling.rs:189-209", "input": [0.0, 1.1, 2.2, 3.3, 4.4, 5.5],
                             "expect": [0.0, 0.2, 0.4, 0.6, 0.8, 1.0], "atol": 3e-7},
    "spectrum_statistics": {"ref": "src/spectrum.rs:654-742",
                            "pairs": [[0.0, 5.0], [50.0, 50.0], [100.0, 100.0], [150.0, 150.0], [200.0, 100.0], [250.0, 20.0],
                                      [300.0, 0.0], [450.0, 200.0], [500.0, 100.0]],
                            "min": [300.0, 0.0], "max": [450.0, 200.0], "range": 200.0, "average": 80.55556, "median": 100.0},
    "fft_adapter_bins": {"ref": "src/fft.rs:141-147", "n": 4, "bins": 3},
}


def oracle_vectors():
    out = []
    cases = [(64, 3, O.WINDOW_HANN, O.SCALING_DIVIDE_BY_N_SQRT), (256, 2, O.WINDOW_HAMMING, O.SCALING_20_TIMES_LOG10),
             (1024, 2, O.WINDOW_BH4, O.SCALING_ZERO_TO_ONE), (4096, 1, O.WINDOW_HANN, O.SCALING_DIVIDE_BY_N_SQRT)]
    for n, frames, win, sc in cases:
        x = O.generate_noise(0, frames * n, 0xC0FFEE + n)
        code, vals, stats, status, _ = O.spectrum_batched(x, frames, n, n, 44100, O.LIMIT_ALL, 0, 0, win, sc)
        assert code == 0 and (status == 0).all()
        idx = sorted({0, 1, 2, n // 8, n // 4, n // 2 - 1, n // 2})
        out.append({"n": n, "frames": frames, "window": int(win), "scaling": int(sc), "seed": 0xC0FFEE + n, "bins": idx,
                    "vals_bits": [[int(np.float32(vals[f][b]).view(np.uint32)) for b in idx] for f in range(frames)],
                    "stats": [{k: (int(stats[k][f]) if "bin" in k else int(np.float32(stats[k][f]).view(np.uint32)))
                               for k in ("min_val", "min_bin", "max_val", "max_bin", "average", "median")} for f in range(frames)]})
    return out


if __name__ == "__main__":
    json.dump(KNOWN, open(os.path.join(HERE, "reference_known_answers.json"), "w"), indent=1)
    json.dump(oracle_vectors(), open(os.path.join(HERE, "oracle_vectors.json"), "w"), indent=1)
    print("written")
