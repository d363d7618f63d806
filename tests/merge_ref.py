"""Host statement of the per-GPU top-k merge (what k_merge_topk computes), used by the CPU multi-rank test."""
import numpy as np


def merge_topk_host(gids, gsc, metric_l2=False):
    """gids, gsc: G x nq x K per-rank results (sorted, -1/NaN padded).  Returns nq x K merged, id-de-duplicated."""
    G, nq, K = gids.shape
    out_i = np.full((nq, K), -1, np.int32)
    out_s = np.full((nq, K), np.nan)
    for q in range(nq):
        ent = {}
        for g in range(G):
            for r in range(K):
                if gids[g, q, r] >= 0:
                    ent[int(gids[g, q, r])] = float(gsc[g, q, r])      # same id => bit-identical score
        order = sorted(ent.items(), key=lambda kv: ((kv[1] if metric_l2 else -kv[1]), kv[0]))[:K]
        for r, (i, s) in enumerate(order):
            out_i[q, r], out_s[q, r] = i, s
    return out_i, out_s
