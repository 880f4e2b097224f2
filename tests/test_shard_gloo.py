"""N>1 host logic on CPU: world_size-2 gloo ranks each build their document shard
(oramacore_b200.sharding), score it with the oracle (global N / avg_len / df), exchange
per-shard products with one all_gather, merge per tests/shard_spec.py, and must reproduce
the unsharded oracle search exactly."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as orc
    import shard_spec
    from oramacore_b200 import synth
    from oramacore_b200.sharding import shard_range, shard_string_index
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n, dim, vocab, B, limit = 6000, 64, 400, 6, 10
    rows = synth.make_vectors(n, dim, seed=1)
    qv, _ = synth.make_vector_queries(rows, B, seed=2)
    data = synth.make_text_corpus(n, vocab, seed=3)
    texts = synth.make_text_queries(vocab, B, seed=4)
    lo, hi = shard_range(n, rank, world)
    sd, gdf = shard_string_index(data, lo, hi)
    assert sd.n_rows == hi - lo and int(sd.row_doc_ids[0]) == lo
    ix = orc.StrIndex(sd, global_df=gdf)
    st = orc.EmbStore(rows[lo:hi], row_doc_ids=np.arange(lo, hi, dtype=np.uint64))
    ok = True
    for mode in (0, 1, 2):
        local = [shard_spec.local_products(orc, ix, st, mode, texts[i], qv[i], limit, 0.0) for i in range(B)]
        gathered = [None] * world
        dist.all_gather_object(gathered, local)          # the one collective of the path
        if rank == 0:
            full_ix, full_st = orc.StrIndex(data), orc.EmbStore(rows)
            sb = orc.SearchBatch(full_ix, full_st)
            for i in range(B):
                sb.add(mode, limit=limit, similarity=0.0, q_vec=qv[i], text=texts[i])
            od, os_, on, oc = sb.run(2)
            for i in range(B):
                docs, scores, count = shard_spec.merge([g[i] for g in gathered], mode, limit, 0, limit)
                ok &= count == int(oc[i])
                ok &= docs == od[i, :on[i]].tolist()
                ok &= bool(np.allclose(scores, os_[i, :on[i]], rtol=0, atol=1e-6))
    # ---- OMC multipliers, offset and the all-tokens threshold go through the same single exchange
    rng0 = np.random.default_rng(7)
    omc_doc = np.sort(rng0.choice(n, size=400, replace=False)).astype(np.uint64)
    omc_mult = rng0.choice([2.0, 3.0, 0.5], size=400).astype(np.float32)
    omc = {int(d): m for d, m in zip(omc_doc, omc_mult)}
    for mode, kw in ((2, dict(limit=5, offset=3, similarity=0.0)), (0, dict(limit=limit, offset=0, threshold=1.0)),
                     (2, dict(limit=limit, offset=0, similarity=0.0, use_omc=True))):
        kw = dict(kw)
        use_omc = kw.pop("use_omc", False)
        thr = kw.get("threshold")
        local = [shard_spec.local_products(orc, ix, st, mode, texts[i], qv[i], kw["limit"], kw.get("similarity", 0.0), threshold=thr)
                 for i in range(B)]
        gathered = [None] * world
        dist.all_gather_object(gathered, local)
        if rank == 0:
            full_ix, full_st = orc.StrIndex(data), orc.EmbStore(rows)
            sb = orc.SearchBatch(full_ix, full_st)
            for i in range(B):
                sb.add(mode, q_vec=qv[i], text=texts[i], omc_doc=omc_doc if use_omc else None,
                       omc_mult=omc_mult if use_omc else None, **kw)
            od, os_, on, oc = sb.run(2)
            for i in range(B):
                docs, scores, count = shard_spec.merge([g[i] for g in gathered], mode, kw["limit"], kw["offset"],
                                                       kw["limit"] + kw["offset"], omc=omc if use_omc else None)
                ok &= count == int(oc[i])
                ok &= docs == od[i, :on[i]].tolist()
                ok &= bool(np.allclose(scores, os_[i, :on[i]], rtol=0, atol=1e-6))
    # ---- with a filter the corpus df is counted per shard and summed with one all_reduce before idf
    import torch
    rng = np.random.default_rng(11)
    allowed = np.sort(rng.choice(n, size=n // 3, replace=False))
    fb = orc.make_filter_bits(allowed.tolist(), n)
    ldf = shard_spec.local_filtered_df(sd, fb, None)
    gdf_f = []
    for a in ldf:
        t = torch.from_numpy(a.copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)     # the df all-reduce (ncclAllReduce on the GPU path)
        gdf_f.append(np.maximum(t.numpy(), 1).astype(np.uint32))
    ix_f = orc.StrIndex(sd, global_df=gdf_f)
    for mode in (0, 2):
        local = [shard_spec.local_products(orc, ix_f, st, mode, texts[i], qv[i], limit, 0.0, filter_bits=fb, filter_nbits=n)
                 for i in range(B)]
        gathered = [None] * world
        dist.all_gather_object(gathered, local)
        if rank == 0:
            full_ix, full_st = orc.StrIndex(data), orc.EmbStore(rows)
            sb = orc.SearchBatch(full_ix, full_st)
            for i in range(B):
                sb.add(mode, limit=limit, similarity=0.0, q_vec=qv[i], text=texts[i], filter_bits=fb, filter_nbits=n)
            od, os_, on, oc = sb.run(2)
            for i in range(B):
                docs, scores, count = shard_spec.merge([g[i] for g in gathered], mode, limit, 0, limit)
                ok &= count == int(oc[i])
                ok &= docs == od[i, :on[i]].tolist()
                ok &= bool(np.allclose(scores, os_[i, :on[i]], rtol=0, atol=1e-6))
    if rank == 0:
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_merge():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok


def test_shard_string_index_partition():
    sys.path.insert(0, ROOT)
    from oramacore_b200 import synth
    from oramacore_b200.sharding import shard_range, shard_string_index
    data = synth.make_text_corpus(5000, 300, seed=9)
    f = data.fields[0]
    tot = 0
    for r in range(4):
        lo, hi = shard_range(5000, r, 4)
        sd, gdf = shard_string_index(data, lo, hi)
        sd.fields[0].validate()
        tot += int(sd.fields[0].term_offsets[-1])
        assert np.array_equal(gdf[0], np.diff(f.term_offsets.astype(np.int64)).astype(np.uint32))
        assert sd.fields[0].post_row.max() < hi - lo and sd.document_count == 5000
    assert tot == int(f.term_offsets[-1])
