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
    # tensor_core = the pipelined one-pass kernel (knn_wide2.cu) where it applies, tensor_core_v1 = the two-pass kernel
    for name, var in (("tensor_core", None), ("tensor_core_v1", "OGMM_KNN_WIDE_V1"), ("fp32_fma", "OGMM_KNN_NO_TENSOR")):
        if name == "fp32_fma" and args.skip_fp32:
            continue
        os.environ.pop("OGMM_KNN_NO_TENSOR", None)
        os.environ.pop("OGMM_KNN_WIDE_V1", None)
        if var:
            os.environ[var] = "1"
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
        res[name + "_idx"] = idx
    os.environ.pop("OGMM_KNN_NO_TENSOR", None)
    os.environ.pop("OGMM_KNN_WIDE_V1", None)
    if "fp32_fma" in res:
        res["speedup"] = res["fp32_fma"]["ms"] / res["tensor_core"]["ms"]
        res["speedup_v1"] = res["fp32_fma"]["ms"] / res["tensor_core_v1"]["ms"]
        res["identical"] = bool(torch.equal(res["tensor_core_idx"], res["fp32_fma_idx"]))
        res["identical_v1"] = bool(torch.equal(res["tensor_core_v1_idx"], res["fp32_fma_idx"]))
    for nm in ("tensor_core", "tensor_core_v1", "fp32_fma"):
        res.pop(nm + "_idx", None)
    fb = og.ops.knn_wide(x, x, args.k)[2]
    res["fallback_queries"] = int(fb)
    out.append(res)
    print(json.dumps(res))
