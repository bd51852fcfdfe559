"""CPU emulation of the histogram-threshold kNN selection planned for round 2.

Per query (3-D, self kNN, k = 20): candidates are visited in order of |x - x_q| (the sorted sweep).  Pass 1 only counts
them into a log-spaced histogram of the squared distance (float bits >> 20: 8 bins per octave); every 8 candidates the
running bound tau = upper edge of the bin holding the k-th smallest count is refreshed and prunes the sweep
(stop when dx^2 > tau).  Pass 2 re-sweeps the slab and collects every candidate with d2 <= tau_final; those are sorted
exactly.  The script reports what the kernel design needs to know: candidates evaluated per pass, size of the
collected set (the register sort network is 32 wide), and how the current kernel's insert count compares.

    python tools/knn_hist_emulation.py [--clouds 8] [--n 1024] [--k 20]
"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ogmm_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--clouds", type=int, default=8)
ap.add_argument("--n", type=int, default=1024)
ap.add_argument("--k", type=int, default=20)
ap.add_argument("--refresh", type=int, default=8)
args = ap.parse_args()

def bin_of(d2):
    return (np.float32(d2).view(np.uint32) >> 20).astype(np.int64)          # sign 0, 8 exponent bits, 3 mantissa bits

def edge_of(b):
    return np.uint32((int(b) + 1) << 20).view(np.float32)                    # smallest float of the next bin

src, _, _, _ = synth.modelnet_batch(0, args.clouds, args.n)
ev1, ev2, coll, ins_now, ok = [], [], [], [], 0
for c in range(args.clouds):
    pts = src[c].T.astype(np.float32)                                          # (N,3)
    ext = pts.max(0) - pts.min(0)
    ax = int(np.argmax(ext))
    for q in range(0, args.n, 7):                                              # a sample of queries
        d2 = ((pts - pts[q]) ** 2).sum(1).astype(np.float32)
        dx2 = ((pts[:, ax] - pts[q, ax]) ** 2).astype(np.float32)
        order = np.argsort(dx2, kind="stable")
        # ---- pass 1: histogram + running bound
        hist = np.zeros(4096, np.int64)
        tau = np.float32(np.inf)
        n1 = 0
        for i, m in enumerate(order):
            if dx2[m] > tau:
                break
            hist[bin_of(d2[m])] += 1
            n1 += 1
            if (i + 1) % args.refresh == 0 and n1 >= args.k:
                cum = np.cumsum(hist)
                tau = edge_of(int(np.searchsorted(cum, args.k)))
        cum = np.cumsum(hist)
        tau = edge_of(int(np.searchsorted(cum, args.k)))
        # ---- pass 2: collect
        slab = order[dx2[order] <= tau]
        got = slab[d2[slab] < tau]
        ev1.append(n1); ev2.append(len(slab)); coll.append(len(got))
        true = np.sort(d2)[args.k - 1]
        ok += int(np.sum(d2 <= true) <= len(got) and np.all(np.isin(np.argsort(d2, kind="stable")[:args.k], got)))
        # ---- today's kernel: sorted inserts with the running k-th best as threshold
        best, cnt = [], 0
        thr = np.float32(np.inf)
        for m in order:
            if dx2[m] > thr:
                break
            if d2[m] <= thr:
                best.append(d2[m]); best.sort(); best = best[:args.k]; cnt += 1
                if len(best) == args.k:
                    thr = best[-1]
        ins_now.append(cnt)
ev1, ev2, coll, ins_now = map(np.array, (ev1, ev2, coll, ins_now))
print(f"queries {len(ev1)}  exact top-{args.k} contained in the collected set: {ok}/{len(ev1)}")
print(f"pass 1 candidates evaluated: mean {ev1.mean():.0f}  p99 {np.percentile(ev1, 99):.0f}")
print(f"pass 2 candidates evaluated: mean {ev2.mean():.0f}  p99 {np.percentile(ev2, 99):.0f}")
print(f"collected set size:          mean {coll.mean():.1f}  p99 {np.percentile(coll, 99):.0f}  max {coll.max()}  (> 32: {(coll > 32).mean() * 100:.2f} %)")
print(f"today's sorted inserts per query (ideal thresholding, no staging lag): mean {ins_now.mean():.0f}  p99 {np.percentile(ins_now, 99):.0f}")
