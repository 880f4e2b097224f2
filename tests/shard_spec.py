"""Python statement of the cross-shard merge (the specification K5 / shard.cuh implements):
inputs are per-shard oracle products, output must equal the unsharded oracle search."""
import numpy as np

f32 = np.float32


def local_filtered_df(sd, filter_bits, n_terms_by_field):
    """Per-shard corpus df under a filter (token_score.rs:262-275 restricted to this shard's
    documents): what every rank feeds the df all-reduce.  One int64 array per field."""
    out = []
    for fi, f in enumerate(sd.fields):
        docs = sd.row_doc_ids[f.post_row] if sd.row_doc_ids is not None else f.post_row.astype(np.uint64)
        ok = ((filter_bits[(docs >> np.uint64(6)).astype(np.int64)] >> (docs & np.uint64(63))) & np.uint64(1)).astype(np.int64)
        cs = np.concatenate([[0], np.cumsum(ok)])
        out.append((cs[f.term_offsets[1:].astype(np.int64)] - cs[f.term_offsets[:-1].astype(np.int64)]).astype(np.int64))
    return out


def local_products(orc, ix, st, mode, text, qv, limit, similarity, threshold=None, filter_bits=None, filter_nbits=0):
    """What one shard contributes for one query (restated with the oracle's pieces)."""
    out = {"count_ft": 0, "max_ft": f32(0), "min_ft": f32(0), "ft": {}, "v": []}
    if mode in (0, 2):
        d, s = orc.fulltext(ix, text, threshold=threshold, filter_bits=filter_bits, filter_nbits=filter_nbits)
        out["count_ft"] = len(d)
        out["ft"] = {int(a): f32(b) for a, b in zip(d, s)}
        if len(s):
            out["max_ft"] = max(f32(0), f32(s.max()))
            out["min_ft"] = min(f32(0), f32(s.min()))
    if mode in (1, 2):
        # local top-`limit` by distance with the rank key, then rescale / threshold (kept prefix)
        dv, sv = orc.vector(st, qv, limit, similarity, filter_bits, filter_nbits)
        for doc, score in zip(dv, sv):
            out["v"].append((int(doc), f32(score)))
    return out


def merge(shards, mode, limit, offset, n_keep, omc=None):
    """shards: list of local_products (non-E5 model: rank key order == score order).
    omc: {doc: multiplier} applied after fusion (search.rs:39-48), before top-N."""
    allv = sorted([v for s in shards for v in s["v"]], key=lambda t: (-t[1], t[0]))[:limit] if mode != 0 else []
    vmap = {}
    for d, sc in allv:
        vmap[d] = f32(vmap.get(d, f32(0)) + sc)
    ftall = {}
    for s in shards:
        ftall.update(s["ft"])
    count = sum(s["count_ft"] for s in shards) + sum(1 for d in vmap if d not in ftall)
    mx = max([f32(0)] + [s["max_ft"] for s in shards] + list(vmap.values()))
    mn = min([f32(0)] + [s["min_ft"] for s in shards] + list(vmap.values()))
    final = {}
    if mode == 0:
        final = dict(ftall)
    elif mode == 1:
        final = dict(vmap)
    else:
        with np.errstate(invalid="ignore", divide="ignore"):
            den = f32(mx - mn)
            for d, v in ftall.items():
                final[d] = f32(f32(v - mn) / den)
            for d, v in vmap.items():
                final[d] = f32(final.get(d, f32(0)) + f32(f32(v - mn) / den))
    if omc:
        final = {d: (f32(v * f32(omc[d])) if d in omc else v) for d, v in final.items()}
    items = sorted([(d, s) for d, s in final.items() if not np.isnan(s)], key=lambda t: (-t[1], t[0]))[:n_keep]
    items = items[offset:offset + limit]
    return [d for d, _ in items], [s for _, s in items], count
