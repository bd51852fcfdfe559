"""Time the feature-space kNN (cfg 4 shape: N=16384, k=20, C in {64,128,256}) on the tensor-core path and on the
FP32 FMA kernel (OGMM_KNN_NO_TENSOR=1).  Run on the GPU box:  python tools/bench_knn_wide.py [--n 16384] [--b 2]"""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ogmm_b200 as og

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=16384)
ap.add_argument("--b", type=int, default=2)
ap.add_argument("--k", type=int, default=20)
ap.add_argument("--cs", type=int, nargs="+", default=[64, 128, 256])
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--skip-fp32", action="store_true")
args = ap.parse_args()
dev = "cuda:0"
out = []
for c in args.cs:
    g = torch.Generator().manual_seed(c)
    x = torch.relu(torch.randn(args.b, args.n, c, generator=g)).to(dev)
    res = {"C": c, "N": args.n, "B": args.b, "k": args.k}
    for name, env in (("tensor_core", None), ("fp32_fma", "1")):
        if env and args.skip_fp32:
            continue
        if env:
            os.environ["OGMM_KNN_NO_TENSOR"] = env
        else:
            os.environ.pop("OGMM_KNN_NO_TENSOR", None)
        for _ in range(2):
            idx = og.knn(x, x, args.k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            idx = og.knn(x, x, args.k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        flops = 2.0 * args.b * args.n * args.n * c
        res[name] = {"ms": ms, "gram_tflops": flops / ms / 1e9, "clouds_per_s": args.b / ms * 1e3}
        res[name + "_idx_sum"] = int(idx.sum())
    os.environ.pop("OGMM_KNN_NO_TENSOR", None)
    if "fp32_fma" in res:
        res["speedup"] = res["fp32_fma"]["ms"] / res["tensor_core"]["ms"]
        res["identical"] = res["tensor_core_idx_sum"] == res["fp32_fma_idx_sum"]
    fb = og.ops.knn_wide(x, x, args.k)[2]
    res["fallback_queries"] = int(fb)
    out.append(res)
    print(json.dumps(res))
