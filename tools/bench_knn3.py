"""A/B of the 3-D kNN kernels on the GPU box: threshold-then-collect selection (knn_select.cu, default) against the
insert-while-sweeping kernel (knn_sweep.cu, OGMM_KNN_SWEEP_INSERT=1) and the exhaustive kernel (OGMM_KNN_EXHAUSTIVE=1).
Results must be bit-identical; prints one JSON line per shape with the timings.

    python tools/bench_knn3.py > gpurun_out/knn3.jsonl
"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ogmm_b200 import ops, synth, _lib

dev = "cuda:0"


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def with_env(name, fn):
    os.environ[name] = "1"
    try:
        return fn()
    finally:
        del os.environ[name]


def select_stats(src, dst, k):
    """Counters of the selection kernel (ogmm_knn3_select_stats), or None outside its range."""
    B, N, _ = src.shape
    M = dst.shape[1]
    if not (256 <= M <= 4096 and N <= 4096 and k <= 24):
        return None
    idx = torch.empty((B, N, k), dtype=torch.int64, device=src.device)
    st = torch.zeros(16, dtype=torch.int32, device=src.device)
    rc = _lib.load().ogmm_knn3_select_stats(src.data_ptr(), *src.stride(), dst.data_ptr(), *dst.stride(), B, N, M, k,
                                            idx.data_ptr(), st.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "ogmm_knn3_select_stats")
    v = st.cpu().tolist()
    w = max(v[1], 1)
    return {"warps": v[1], "redo_frac": v[0] / w, "steps_per_warp": v[2] / w, "merges_per_warp": v[3] / w,
            "insert_rounds_per_warp": v[4] / w, "max_groups_per_warp": v[5] / w, "tie_frac": v[7] / w,
            "few_groups_frac": v[8] / w}


def case(name, src, dst, k, edge, reps=20):
    run = lambda: ops.knn_graph(src, dst, k, want_edge=edge, want_dist=True)
    new = run()
    old = with_env("OGMM_KNN_SWEEP_INSERT", run)
    exh = with_env("OGMM_KNN_EXHAUSTIVE", run)
    same_old = all(torch.equal(a, b) for a, b in zip(new, old) if a is not None)
    same_exh = all(torch.equal(a, b) for a, b in zip(new, exh) if a is not None)
    t_new = timed(run, reps)
    t_old = with_env("OGMM_KNN_SWEEP_INSERT", lambda: timed(run, reps))
    row = {"case": name, "B": src.shape[0], "N": src.shape[1], "M": dst.shape[1], "k": k, "edge": edge,
           "identical_to_sweep_insert": same_old, "identical_to_exhaustive": same_exh,
           "select_ms": t_new, "sweep_insert_ms": t_old, "speedup": t_old / t_new, "stats": select_stats(src, dst, k)}
    print(json.dumps(row), flush=True)
    return row


if __name__ == "__main__":
    _lib.load()
    h = synth.hot_path_inputs(0, 256, 1024, 8, tile=32)
    src = torch.from_numpy(h["src"]).to(dev).transpose(1, 2)            # (256,1024,3) view of (B,3,N)
    tgt = torch.from_numpy(h["tgt"]).to(dev).transpose(1, 2)
    case("bench clouds src k=20 +edge", src, src, 20, True)
    case("bench clouds tgt k=20 +edge", tgt, tgt, 20, True)
    case("bench clouds k=5 (PositionEncoding)", src, src, 5, False)
    g = torch.Generator().manual_seed(1)
    anchors = src[:, :128].contiguous()
    case("anchors 128 vs 1024, k=1 (get_local_corrs)", anchors, src, 1, False)
    u = (torch.rand(64, 1024, 3, generator=g) * 2 - 1).to(dev)
    case("uniform cube 1024 k=20", u, u, 20, True)
    u2 = (torch.rand(16, 4096, 3, generator=g) * 2 - 1).to(dev)
    case("uniform cube 4096 k=16", u2, u2, 16, True, reps=5)
    r = (torch.rand(8, 717, 3, generator=g)).to(dev)
    case("ragged 717 k=20", r, r, 20, True)
    r2 = (torch.rand(4, 300, 3, generator=g)).to(dev)
    r3 = (torch.rand(4, 1500, 3, generator=g)).to(dev)
    case("two clouds 300 vs 1500 k=8", r2, r3, 8, False)
    icl, _, _, _ = synth.icl_nuim_batch(0, 32, 1024)
    ic = torch.from_numpy(icl).to(dev).transpose(1, 2)
    case("ICL-NUIM-shape (planes) k=20", ic, ic, 20, True)
    plane = torch.rand(8, 1024, 3, generator=g)
    plane[:, :, 0] = 0.25                                                # every point on one plane orthogonal to x
    plane = plane.to(dev)
    case("degenerate plane k=20", plane, plane, 20, True)
    dup = torch.rand(4, 256, 3, generator=g).repeat(1, 4, 1).to(dev)     # every point four times
    case("4x duplicated points k=20", dup, dup, 20, True)
    same = torch.full((2, 512, 3), 0.5, device=dev)
    case("all points identical k=20", same, same, 20, True)
