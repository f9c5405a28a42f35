"""Small-K contraction paths side by side (north star: "warp-shuffle FFMA for small K, with the choice justified by ncu
tensor-pipe / FP32-pipe utilisation and achieved HBM GB/s"): per-iteration time of VB_aspect on the FP64 tensor pipe
(DMMA, the AUTO choice for K <= 32), the FP32 FFMA kernel, the FP64 SIMT kernel (Kp > 128 only in production; forced
here with NBMF_FP64_KERNEL=simt in a child process) and -- where K allows -- the tcgen05 path.
    python tools/smallk_bench.py [quick] [shapes=i,j,...] [iters=n]      (shape indices into SHAPES; for ncu captures)"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = [(200, 300, 3, "C1"), (100, 5429, 10, "C2 shape"), (784, 60000, 16, "C3 shape (all of MNIST)"), (8192, 8192, 16, ""),
          (16384, 16384, 32, ""), (8192, 8192, 64, ""), (16384, 16384, 128, "")]


def run(F, N, K, prec, iters):
    from nbmf_b200 import Engine
    e = Engine(precision=prec)
    e.generate_synthetic(F, N, min(K, 8), 0.5, 0.5, 3)
    e.init_random_labels(K, 4)
    e.vb_begin(K, 1.0, 1.0, 1.0)
    for _ in range(3):
        e.vb_step(True)
    e.vb_end(0, want_outputs=False)
    e.vb_begin(K, 1.0, 1.0, 1.0)
    for _ in range(iters):
        e.vb_step(True)
    e.vb_end(0, want_outputs=False)
    t = e.timing()
    e.close()
    return t


def main():
    from nbmf_b200 import _lib
    simt = os.environ.get("NBMF_FP64_KERNEL") == "simt"
    out = []
    pick = [a.split("=")[1] for a in sys.argv if a.startswith("shapes=")]
    shapes = [SHAPES[int(i)] for i in pick[0].split(",")] if pick else SHAPES
    fixed = [int(a.split("=")[1]) for a in sys.argv if a.startswith("iters=")]
    for F, N, K, tag in shapes:
        iters = fixed[0] if fixed else (20 if F * N * K < 1e10 else 5)
        modes = [("fp64 simt", _lib.PREC_FP64)] if simt else [("fp64 dmma", _lib.PREC_FP64), ("fp32 ffma", _lib.PREC_FP32)]
        if not simt and K > 32:
            modes.append(("tcgen05 hi+lo8", _lib.PREC_TENSOR_LO8))
        for name, prec in modes:
            t = run(F, N, K, prec, iters)
            ms = t["total_ms"] / iters
            rec = dict(F=F, N=N, K=K, tag=tag, kernel=name, ms_per_iter=round(ms, 4), pass_rows_ms=round(t["pass_rows_ms"], 4),
                       pass_cols_ms=round(t["pass_cols_ms"], 4), params_ms=round(t["params_ms"], 4),
                       entries_K_per_s=F * N * K / (ms * 1e-3), counted_tflops=12.0 * F * N * K / ((t["pass_rows_ms"] + t["pass_cols_ms"]) * 1e-3) / 1e12)
            out.append(rec)
            print(json.dumps(rec), flush=True)
    if not simt and "quick" not in sys.argv:
        env = dict(os.environ, NBMF_FP64_KERNEL="simt")
        subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, check=False)


if __name__ == "__main__":
    main()
