"""CPU model of the scan/flush control flow of knn_scan_kernel for ONE warp (development aid):
counts flush rounds, buffered groups, survivors and merge-path usage so that buffer sizes and
thresholds can be tuned offline.  Uniform random points, float64 arithmetic (counts only)."""
import sys
import numpy as np

def sim(P=16384, K=16, Q=3, RS=2048, BCAP=16, CHUNK=4, SCAP=16, seed=0, G=4, ins_thresh=5):
    rng = np.random.default_rng(seed)
    pts = rng.random((P, 3))
    nq = 32 * Q
    qs = pts[rng.choice(P, nq, replace=False)]
    d2 = ((qs[:, None, :] - pts[None, :, :]) ** 2).sum(-1)  # (nq, P)
    lists = np.full((nq, K), np.inf)
    T = np.full(nq, np.inf)
    buf = [[] for _ in range(nq)]
    stats = dict(rounds=0, forced=0, overflow=0, fill_iters=0, fill_entries=0, merges_ins=0, ins_steps=0,
                 merges_sort=0, survivors=0, inserted=0, subrounds=0)
    ngroups_total = P // G

    def flush():
        stats["rounds"] += 1
        for t in range(Q):
            sl = slice(t * 32, (t + 1) * 32)
            cnts = np.array([len(b) for b in buf[sl]])
            if cnts.max() == 0:
                continue
            c = np.zeros(32, int)
            while (c < cnts).any():
                stats["subrounds"] += 1
                ns = np.zeros(32, int)
                iters = 0
                surv = [[] for _ in range(32)]
                active = (c < cnts) & (ns <= SCAP - G)
                while active.any():
                    iters += 1
                    for l in np.nonzero(active)[0]:
                        q = t * 32 + l
                        g = buf[q][c[l]]
                        c[l] += 1
                        stats["fill_entries"] += 1
                        dd = d2[q, g * G:(g + 1) * G]
                        for v in dd[dd <= lists[q, K - 1]]:
                            surv[l].append(v)
                            ns[l] += 1
                    active = (c < cnts) & (ns <= SCAP - G)
                stats["fill_iters"] += iters
                nsm = ns.max()
                stats["survivors"] += ns.sum()
                if nsm > 0:
                    if nsm <= ins_thresh:
                        stats["merges_ins"] += 1
                        stats["ins_steps"] += nsm
                    else:
                        stats["merges_sort"] += 1
                    for l in range(32):
                        if surv[l]:
                            q = t * 32 + l
                            before = lists[q].copy()
                            allv = np.sort(np.concatenate([lists[q], np.array(surv[l])]))[:K]
                            stats["inserted"] += int((~np.isin(allv, before)).sum())
                            lists[q] = allv
            for q in range(t * 32, (t + 1) * 32):
                buf[q] = []
                T[q] = lists[q, K - 1]

    for tile0 in range(0, P, RS):
        ng = min(RS, P - tile0) // G
        for g0 in range(0, ng, CHUNK):
            for g in range(g0, g0 + CHUNK):
                gg = tile0 // G + g
                m = d2[:, gg * G:(gg + 1) * G].min(1)
                for q in np.nonzero(m <= T)[0]:
                    buf[q].append(gg)
            if max(len(b) for b in buf) > BCAP - CHUNK:
                stats["overflow"] += 1
                flush()
        stats["forced"] += 1
        flush()
    return stats

if __name__ == "__main__":
    kw = {}
    for a in sys.argv[1:]:
        k, v = a.split("=")
        kw[k] = int(v)
    st = sim(**kw)
    Q = kw.get("Q", 3)
    nq = 32 * Q
    print(kw)
    for k, v in st.items():
        print(f"  {k}: {v}   per-query {v / nq:.1f}" if k in ("fill_entries", "survivors", "inserted") else f"  {k}: {v}")
    # rough warp-instruction model
    fill = st["fill_iters"] * 50
    merge = st["merges_ins"] * 48 + st["ins_steps"] * 105 + st["merges_sort"] * 700
    scan = (kw.get("P", 16384) // 4) * (Q * 11 + 6)
    print(f"  model warp-instr: scan {scan}  fill {fill}  merge {merge}  flush/scan {(fill + merge) / scan:.2f}")
