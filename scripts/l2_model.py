"""Offline model of the L2 hit rate of the walk kernel's neighbour loads (no GPU needed).

A walker sits at v with probability ~ PageRank(v) (alpha = 0.2, near-uniform restart: the residue-weighted
starts of a FORA query are spread wide) and reads ONE random element of v's adjacency list, so a 32-byte
sector of the column array is referenced with probability sum_{elements in sector} pr(v)/d_out(v).
Under the independent-reference model this script compares
  * LRU (Che's approximation) -- what the hardware does by default,
  * a static cache of the most popular sectors -- what evict_last/evict_first hints approximate,
for the in-degree vertex order the engine uses and for an order by pr(v)/d_out(v).
"""
import sys
import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from helpers import synth_edges  # noqa: E402


def pagerank(n, src, dst, deg, alpha=0.2, iters=30):
    pr = np.full(n, 1.0 / n)
    for _ in range(iters):
        share = np.where(deg > 0, pr / np.maximum(deg, 1), 0.0)
        nxt = np.bincount(dst, weights=share[src], minlength=n) * (1 - alpha)
        dang = pr[deg == 0].sum() * (1 - alpha)
        nxt += (alpha + dang) / n
        pr = nxt / nxt.sum()
    return pr


def che_hit(p, cap):
    """LRU hit rate for reference probabilities p (sum 1) and capacity cap items."""
    lo, hi = 1.0, 1e12
    for _ in range(80):
        T = np.sqrt(lo * hi)
        occ = (1.0 - np.exp(-p * T)).sum()
        if occ > cap:
            hi = T
        else:
            lo = T
    return float((p * (1.0 - np.exp(-p * lo))).sum())


def sector_popularity(order, deg, w):
    """column array laid out in `order`; returns per-sector reference probability."""
    d = deg[order]
    per_elem = np.where(d > 0, w[order] / np.maximum(d, 1), 0.0)
    beg = np.concatenate(([0], np.cumsum(d)))[:-1]
    m = int(d.sum())
    elem = np.repeat(per_elem, d)
    nsec = (m + 7) // 8
    pad = nsec * 8 - m
    if pad:
        elem = np.concatenate((elem, np.zeros(pad)))
    del beg
    return elem.reshape(nsec, 8).sum(1)


def main():
    n, m = (4847571, 68993773) if len(sys.argv) < 2 else (int(sys.argv[1]), int(sys.argv[2]))
    src, dst = synth_edges(n, m, 42)
    deg = np.bincount(src, minlength=n)
    indeg = np.bincount(dst, minlength=n)
    pr = pagerank(n, src, dst, deg)
    hop = np.where(deg > 0, pr, 0.0)
    hop /= hop.sum()
    orders = {
        "in-degree (engine)": np.argsort(-indeg, kind="stable"),
        "pr/d_out": np.argsort(-(hop / np.maximum(deg, 1)), kind="stable"),
    }
    for name, order in orders.items():
        p = sector_popularity(order, deg, hop)
        p /= p.sum()
        ps = np.sort(p)[::-1]
        cum = np.cumsum(ps)
        pre = np.cumsum(p)
        print("order:", name, " sectors:", len(p))
        for mb in (8, 16, 24, 32, 48, 64, 96):
            cap = mb * (1 << 20) // 32
            print("  %3d MB: LRU(Che) %.3f   static-best-sectors %.3f   static-prefix %.3f" % (
                mb, che_hit(p, cap), cum[min(cap, len(cum)) - 1], pre[min(cap, len(pre)) - 1]))
    # row-offset loads for comparison: probability pr(v) over 4-byte entries
    for name, order in orders.items():
        q = pr[order]
        nsec = (n + 7) // 8
        q = np.concatenate((q, np.zeros(nsec * 8 - n))).reshape(nsec, 8).sum(1)
        cum = np.cumsum(q)
        print("row offsets, order %s: prefix 4 MB %.3f  8 MB %.3f  19 MB %.3f" % (
            name, cum[(4 << 20) // 32 - 1], cum[(8 << 20) // 32 - 1], cum[-1]))


if __name__ == "__main__":
    main()
