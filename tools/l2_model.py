"""Row-granular LRU model of the L2 for the backward CSC pass (analysis only, CPU): how many of the t-stream row
gathers would hit for (a) the plain column order, (b) target-range blocking, (c) pinning the rows of hub targets.
    python tools/l2_model.py [arxiv|mag]
"""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, synth_graph  # noqa: E402


def lru_hits(stream, capacity):
    cache, hits = OrderedDict(), 0
    for r in stream:
        if r in cache:
            cache.move_to_end(r)
            hits += 1
        else:
            cache[r] = None
            if len(cache) > capacity:
                cache.popitem(last=False)
    return hits


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "arxiv"
    w = WORKLOADS[wl]
    n, ei = synth_graph(wl, 0)
    loops = torch.arange(n)
    src = torch.cat([ei[0], loops]).numpy()
    dst = torch.cat([ei[1], loops]).numpy()
    order = np.lexsort((dst, src))                       # CSC: by source column, targets ascending inside a column
    targets = dst[order]
    a = len(w["aggrs"])
    sym = int("symnorm" in w["aggrs"]); lin = int(any(x in ("sum", "mean", "var", "std") for x in w["aggrs"])); sq = int(any(x in ("var", "std") for x in w["aggrs"]))
    row_bytes = (sym + lin + sq) * w["bases"] * (w["f_out"] // w["heads"]) * 4
    print(f"{wl}: {n} rows of {row_bytes} B ({n * row_bytes / 1e6:.0f} MB), {targets.size} gathers")
    for l2_mb in (126, 63):
        cap = int(l2_mb * 1e6 * 0.8 / row_bytes)         # 80 % of the capacity usable for this table
        base = lru_hits(targets.tolist(), cap) / targets.size
        line = f"  L2 {l2_mb:3d} MB -> {cap} rows: plain order hit rate {base:.1%}"
        for parts in (2, 3, 4):
            bounds = np.linspace(0, n, parts + 1).astype(np.int64)
            hits = 0
            for p in range(parts):
                sel = targets[(targets >= bounds[p]) & (targets < bounds[p + 1])]
                hits += lru_hits(sel.tolist(), cap)
            line += f"; {parts} target ranges {hits / targets.size:.1%}"
        deg = np.bincount(targets, minlength=n)
        hot = np.argsort(-deg)[:cap // 2]                # pin the hottest rows in half of the capacity, LRU in the rest
        is_hot = np.zeros(n, dtype=bool); is_hot[hot] = True
        cold_stream = targets[~is_hot[targets]]
        pinned = (is_hot[targets].sum() - hot.size + lru_hits(cold_stream.tolist(), cap - hot.size)) / targets.size
        print(line + f"; hub pinning {pinned:.1%}")


if __name__ == "__main__":
    main()
