"""torchrun worker for the document-sharded path: every rank loads its shard, one NCCL
all-gather + on-device merge per batch, and every rank must hold the unsharded oracle answer.
Run: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/sharded_worker.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import oracle as orc
    import oramacore_b200 as ob
    from helpers import assert_topk_equal
    from oramacore_b200 import synth
    from oramacore_b200.sharding import shard_range, shard_string_index

    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ctx = ob.Context(lr)
    uid = [ob.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(world, rank, uid[0])
    if os.environ.get("OC_SHARD_P2P", "1") != "0":     # the records travel by direct NVLink stores (else ncclAllGather)
        def _ag(blob):
            out = [None] * world
            dist.all_gather_object(out, blob)
            return out
        ctx.comm_enable_p2p(_ag)

    n, dim, vocab, B = 60000, 384, 3000, 12
    rows = synth.make_vectors(n, dim, seed=41)
    rows[1000:7000] = rows[5]                  # 6000 exact duplicates in rank 0's shard: more ties than the re-score budget
    qv, _ = synth.make_vector_queries(rows, B, seed=42)
    qv[3] = rows[5] * 2.0                      # -> rank 0's tensor-core scan flags query 3; EVERY rank must re-run the tail
    data = synth.make_text_corpus(n, vocab, seed=43)
    texts = synth.make_text_queries(vocab, B, seed=44)
    lo, hi = shard_range(n, rank, world)
    sd, gdf = shard_string_index(data, lo, hi)
    emb = ob.EmbeddingFieldStorage(ctx, "BGESmall")
    emb.insert_batch(np.arange(lo, hi, dtype=np.uint64), rows[lo:hi])
    strs = ob.StringFieldStorage(ctx, sd, global_df=gdf)
    rng = np.random.default_rng(5)
    omc_doc = np.sort(rng.choice(n, size=3000, replace=False)).astype(np.uint64)
    omc_mult = rng.choice([2.0, 3.0, 0.5], size=3000).astype(np.float32)

    ix, st = orc.StrIndex(data), orc.EmbStore(rows)
    cases = [("hybrid", 2, dict(limit=10, similarity=0.0)), ("hybrid", 2, dict(limit=10, similarity=0.7)),
             ("hybrid", 2, dict(limit=5, offset=4, similarity=0.0, threshold=1.0)),
             ("fulltext", 0, dict(limit=10)), ("vector", 1, dict(limit=10, similarity=0.0)),
             ("hybrid", 2, dict(limit=10, similarity=0.0, omc=True))]
    for name, mode, kw in cases:
        kw = dict(kw)
        omc = kw.pop("omc", False)
        extra = dict(omc_doc_ids=omc_doc, omc_mult=omc_mult) if omc else {}
        hits = ob.search(ctx, emb if mode else None, strs if mode != 1 else None, name, texts=texts if mode != 1 else None,
                         q_vecs=qv if mode else None, sharded=True, **kw, **extra)
        sb = orc.SearchBatch(ix, st)
        for i in range(B):
            sb.add(mode, q_vec=qv[i], text=texts[i], omc_doc=omc_doc if omc else None, omc_mult=omc_mult if omc else None, **kw)
        od, os_, on, oc = sb.run(4)
        for i, h in enumerate(hits):           # every rank holds the global answer (all-gather, not gather)
            assert h.count == int(oc[i]), (name, rank, i, h.count, int(oc[i]))
            assert_topk_equal(h.doc_ids, h.scores, od[i, :on[i]], os_[i, :on[i]], atol=1e-5)
        t = ctx.last_timing()
        if mode and rank == 0:
            assert t["scan_tensor_core"] == 1 and t["scan_unproven"] >= 1, t      # the flagged query was re-run exactly
        if mode:
            assert t["rerun_ms"] > 0.0, (rank, t)                                  # ... and every rank re-entered the collective
        if rank == 0:
            print(f"sharded {name} {kw} omc={omc}: ok on {world} ranks; comm_ms={t['comm_ms']:.3f} device_ms={t['device_ms']:.3f} rerun_ms={t['rerun_ms']:.3f}")
    # ---- corpus df counted on device and summed across ranks (one ncclAllReduce): filter, then tombstones
    allowed = np.sort(rng.choice(n, size=n // 3, replace=False))
    fb = orc.make_filter_bits(allowed.tolist(), n)
    gone = [int(d) for d in rng.choice(n, size=40, replace=False)]
    mine = np.array([d for d in gone if lo <= d < hi], dtype=np.uint64)
    alive = orc.make_filter_bits([d for d in range(n) if d not in set(gone)], n)
    for name, mode, kw, filt, tomb in [("hybrid", 2, dict(limit=10, similarity=0.0), fb, False),
                                       ("fulltext", 0, dict(limit=10), fb, False),
                                       ("fulltext", 0, dict(limit=10), None, True)]:
        if tomb and len(mine):
            strs.delete(mine)                    # only the owning shard tombstones the doc; the flag is global
        extra = dict(filtered_doc_ids=filt, filter_nbits=n) if filt is not None else {}
        hits = ob.search(ctx, emb if mode else None, strs, name, texts=texts, q_vecs=qv if mode else None, sharded=True,
                         shard_tombstones=tomb, **kw, **extra)
        ofb = filt if filt is not None else alive
        sb = orc.SearchBatch(ix, st)
        for i in range(B):
            sb.add(mode, q_vec=qv[i], text=texts[i], filter_bits=ofb, filter_nbits=n, **kw)
        od, os_, on, oc = sb.run(4)
        for i, h in enumerate(hits):
            assert h.count == int(oc[i]), (name, rank, i, h.count, int(oc[i]))
            assert_topk_equal(h.doc_ids, h.scores, od[i, :on[i]], os_[i, :on[i]], atol=1e-5)
        if rank == 0:
            print(f"sharded {name} filter={filt is not None} tombstones={tomb}: ok on {world} ranks (df all-reduce)")
    # ---- a shard without the replicated df tables (what a commit leaves behind): corpus-wide N / avg length are
    # supplied by the caller (oc_str_set_global) and df is counted across ranks (OC_SHARD_COUNT_DF)
    strs2 = ob.StringFieldStorage(ctx, sd)                     # no global_df
    strs2.set_global(data.document_count, [data.fields[0].avg_field_len])
    try:
        ob.search(ctx, None, strs2, "fulltext", texts=texts, sharded=True, limit=10)
        raise AssertionError("a shard-local posting-list length must never be taken for the corpus df")
    except ob.OcError as e:
        assert "OC_SHARD_COUNT_DF" in str(e)
    hits = ob.search(ctx, emb, strs2, "hybrid", texts=texts, q_vecs=qv, sharded=True, shard_count_df=True, limit=10, similarity=0.0)
    sb = orc.SearchBatch(ix, st)
    for i in range(B):
        sb.add(2, q_vec=qv[i], text=texts[i], limit=10, similarity=0.0)
    od, os_, on, oc = sb.run(4)
    for i, h in enumerate(hits):
        assert h.count == int(oc[i]), ("count_df", rank, i, h.count, int(oc[i]))
        assert_topk_equal(h.doc_ids, h.scores, od[i, :on[i]], os_[i, :on[i]], atol=1e-5)
    if rank == 0:
        print(f"sharded hybrid with counted df (no replicated tables): ok on {world} ranks")
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("SHARDED_OK")


if __name__ == "__main__":
    main()
