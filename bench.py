#!/usr/bin/env python
"""bench.py — the hot path of BASELINE.json's metric on synthetic corpora of the named shape.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload h1|v1|t1] [--batch B]
    python bench.py --impl reference ...      # the CPU restatement (oracle) on the host cores

A "step" = one pass of the hot path over one batch of B queries through the C ABI
(oc_search): hybrid = embedding scan + BM25 posting scorer + fusion/top-k.
  value : whole-job QPS with inputs resident in HBM — B*K / sum of the library's own
          CUDA-event device time (H2D of queries .. last kernel), max over ranks.
  e2e   : QPS through the public call with HOST buffers (H2D + kernels + D2H inside),
          K calls bracketed by barrier + device synchronize, max over ranks.
Under torchrun (N>1) the corpus is sharded by document across ranks (strong scaling: the
named corpus is fixed); one NCCL all-gather of per-shard top-k per batch, merged on device.
The matrix (3.07 GB at 1M x 768) is far larger than L2 (126 MB), so no L2 flush is needed
between iterations.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[3] (the config the hybrid-QPS metric is quoted on), configs[1], configs[2]
    "h1": dict(mode="hybrid", n_docs=1_000_000, dim=768, vocab=200_000, batch=256,
               desc="hybrid vector+BM25, 1M docs x 768-d fp32, cosine+BM25F top-10 (BASELINE configs[3])"),
    "v1": dict(mode="vector", n_docs=1_000_000, dim=768, vocab=0, batch=1,
               desc="1M x 768-d fp32 embeddings, cosine top-10, batch=1 (BASELINE configs[1])"),
    "t1": dict(mode="fulltext", n_docs=10_000_000, dim=0, vocab=1_000_000, batch=256,
               desc="BM25 fulltext, 10M synthetic docs (Zipf), batch=256 (BASELINE configs[2])"),
    "v2": dict(mode="vector", n_docs=10_000_000, dim=1024, vocab=0, batch=1024, dtype="bf16",
               desc="10M x 1024-d bf16 embeddings, cosine top-10, batch=1024 (BASELINE configs[4])"),
}


# oc_timing.scan_variant (include/oramacore_b200.h OC_SCAN_*) -> (kernel, description)
SCAN_VARIANTS = {
    0: ("emb_scan_kernel", "exact fp32 sweep"),
    1: ("emb_gemm_kernel", "tcgen05 kind::tf32 on the fp32 rows + exact fp32 re-score"),
    2: ("emb_gemm_pair_kernel", "tcgen05 cta_group::2 kind::tf32 on the fp32 rows + exact fp32 re-score"),
    3: ("emb_gemm_cvt_kernel", "tcgen05 cta_group::2 kind::f16, fp32 rows streamed once and rounded to bf16 in the SM, + exact fp32 re-score"),
    4: ("emb_gemm_kernel", "tcgen05 kind::f16 on the bf16 rows + exact fp32 re-score"),
    5: ("emb_gemm_pair_kernel", "tcgen05 cta_group::2 kind::f16 on the bf16 rows + exact fp32 re-score"),
}

def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="h1", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--n-docs", type=int, default=0, help="override corpus size (debug)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_workload(w, n_docs, batch, rank, world):
    """Synthetic corpus of the named shape; shard = contiguous doc-row range."""
    from oramacore_b200 import synth
    lo, hi = (n_docs * rank) // world, (n_docs * (rank + 1)) // world
    out = dict(lo=lo, hi=hi)
    if w["dim"] and w.get("dtype") == "bf16":
        # too large to hold in fp32 on the host: rows are generated chunk by chunk at load time (see main);
        # queries are planted on rows of the first chunk
        first = synth.make_vectors(min(n_docs, 1 << 18), w["dim"])
        qv, planted = synth.make_vector_queries(first, batch)
        out.update(rows=None, qv=qv, planted=planted, chunked=True)
    elif w["dim"]:
        rows = synth.make_vectors(n_docs, w["dim"])          # deterministic: every rank draws the same stream
        qv, planted = synth.make_vector_queries(rows, batch)
        out.update(rows=rows[lo:hi], rows_all=rows if world == 1 else None, qv=qv, planted=planted)
        if world > 1:
            del rows
    if w["vocab"]:
        data = synth.make_text_corpus(n_docs, w["vocab"])
        out.update(data_all=data, texts=synth.make_text_queries(w["vocab"], batch))
    return out


def run_reference(args, w, batch, n_docs):
    """--impl reference: the reference's CPU algorithm (oracle port; the Rust reference cannot be
    built here) on all host cores, each step a bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    orc.build()
    wl = make_workload(w, n_docs, batch, 0, 1)
    cores = os.cpu_count() or 1
    ix = orc.StrIndex(wl["data_all"]) if w["vocab"] else None
    st = orc.EmbStore(wl["rows"]) if w["dim"] else None
    mode = {"fulltext": 0, "vector": 1, "hybrid": 2}[w["mode"]]
    sample = min(batch, max(cores, 8))
    threads = cores

    def one_step(k):
        sb = orc.SearchBatch(ix, st)
        for i in range(sample):
            j = (k * sample + i) % batch
            sb.add(mode, limit=10, similarity=0.0, q_vec=wl["qv"][j] if w["dim"] else None,
                   text=wl["texts"][j] if w["vocab"] else None)
        t0 = time.perf_counter()
        sb.run(threads)
        return time.perf_counter() - t0

    # bound the run to a few minutes: a step is one query per thread; when a full-width step is too
    # long for steps+warmup of them (the scan is DRAM-bound on the host, so time ~ queries in flight),
    # shrink the per-step sample and the thread count together and report the threads actually used
    t_probe = one_step(0)
    budget = 150.0
    n_steps_total = args.steps + max(args.warmup - 1, 0)
    if t_probe * n_steps_total > budget:
        scale = budget / (t_probe * n_steps_total)
        sample = threads = max(8, min(sample, int(sample * scale)))
    for k in range(1, args.warmup):
        one_step(k)
    times = [one_step(k) for k in range(args.steps)]
    total = sum(times)
    qps = sample * args.steps / total
    cores = threads
    line = {"impl": "reference",
            "metric": "hybrid_search_qps_at_recall10_ge_0.99_1Mx768" if args.workload == "h1" else f"{w['mode']}_search_qps",
            "value": qps, "unit": "queries/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "batch": batch, "n_docs": n_docs, "dim": w["dim"], "vocab": w["vocab"],
                       "limit": 10, "similarity": 0.0, "sample_queries_per_step": sample},
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} queries/step x {args.steps} steps, one query per thread"},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    args = parse()
    w = dict(WORKLOADS[args.workload])
    batch = args.batch or w["batch"]
    n_docs = args.n_docs or w["n_docs"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            run_reference(args, w, batch, n_docs)
        return

    import torch
    import torch.distributed as dist
    import oramacore_b200 as ob

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = ob.Context(local_rank)
    if world > 1:
        uid = [ob.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(world, rank, uid[0])

    wl = make_workload(w, n_docs, batch, rank, world)
    lo, hi = wl["lo"], wl["hi"]
    emb = strs = None
    if w["dim"]:
        emb = ob.EmbeddingFieldStorage(ctx, dim=w["dim"], model="BGEBase" if w["dim"] == 768 else "BGELarge",
                                       dtype=w.get("dtype", "f32"))
        emb.reserve(hi - lo)
        ids = np.arange(lo, hi, dtype=np.uint64)
        if wl.get("chunked"):
            from oramacore_b200 import synth
            CH = 1 << 18
            for c0 in range(0, n_docs, CH):        # chunk c uses seed SEED+c (chunk 0 == the planted chunk)
                c1 = min(n_docs, c0 + CH)
                a, b = max(c0, lo), min(c1, hi)
                if a >= b:
                    continue
                chunk = synth.make_vectors(c1 - c0, w["dim"], seed=synth.SEED_VECTORS + (c0 // CH) * (c0 > 0))
                emb.insert_batch(np.arange(a, b, dtype=np.uint64), chunk[a - c0:b - c0])
        else:
            for i in range(0, hi - lo, 1 << 18):
                emb.insert_batch(ids[i:i + (1 << 18)], wl["rows"][i:i + (1 << 18)])
    if w["vocab"]:
        if world == 1:
            strs = ob.StringFieldStorage(ctx, wl["data_all"])
        else:
            from oramacore_b200.sharding import shard_string_index
            sd, gdf = shard_string_index(wl["data_all"], lo, hi)
            strs = ob.StringFieldStorage(ctx, sd, global_df=gdf)
    mode = {"fulltext": ob.MODE_FULLTEXT, "vector": ob.MODE_VECTOR, "hybrid": ob.MODE_HYBRID}[w["mode"]]
    tsc = ob.TokenScoreContext(ctx, emb, strs)
    params = ob.TokenScoreParams(mode=mode, limit_hint=10, similarity=0.0, sharded=world > 1)
    texts = wl.get("texts")
    qv = wl.get("qv")
    # the step's inputs as they sit in host memory: resolved term ids (packed CSR) + query vectors
    packed = ob.TextQueryBatch(texts) if texts is not None else None
    qv_host = None
    if qv is not None:          # the step's query vectors sit in pinned host memory (DMA'd by oc_search)
        qv_host = ob.pinned_empty(qv.shape, np.float32)
        qv_host[...] = qv

    def step():
        return tsc.execute_batch_arrays(params, packed, qv_host)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        raw = step()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms = scan_ms = bm_ms = fuse_ms = comm_ms = sweep_ms = 0.0
    scan_bytes = scan_launches = postings = h2d = d2h = unproven = tensor_core = variant = 0
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        raw = step()
        t = ctx.last_timing()
        dev_ms += t["device_ms"]; scan_ms += t["scan_ms"]; bm_ms += t["bm25_ms"]; fuse_ms += t["fuse_ms"]
        comm_ms += t["comm_ms"]; sweep_ms += t["scan_sweep_ms"]; scan_bytes += t["scan_bytes"]; scan_launches += t["scan_launches"]
        postings += t["bm25_postings"]; h2d, d2h = t["h2d_bytes"], t["d2h_bytes"]
        unproven += t["scan_unproven"]; tensor_core = max(tensor_core, t["scan_tensor_core"]); variant = max(variant, t["scan_variant"])
    sync_all()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count() - launches0
    hits = [ob.SearchHits(raw[0][i, :raw[2][i]].copy(), raw[1][i, :raw[2][i]].copy(), int(raw[3][i])) for i in range(batch)]

    if world > 1:
        red = torch.tensor([dev_ms, wall * 1e3, scan_ms, bm_ms, fuse_ms, comm_ms, sweep_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms, scan_ms, bm_ms, fuse_ms, comm_ms, sweep_ms = red.tolist()
        tot = torch.tensor([float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        launches = int(tot.item())
    else:
        wall_ms = wall * 1e3
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    K, B = args.steps, batch
    value = B * K / (dev_ms * 1e-3)
    e2e = B * K / (wall_ms * 1e-3)
    peak, peak_src = peaks()
    line = {
        "metric": "hybrid_search_qps_at_recall10_ge_0.99_1Mx768" if args.workload == "h1" else f"{w['mode']}_search_qps",
        "value": value, "unit": "queries/s", "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16 storage, f32 arithmetic" if w.get("dtype") == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": w["desc"], "batch": B, "n_docs": n_docs, "dim": w["dim"], "vocab": w["vocab"],
                   "limit": 10, "similarity": 0.0, "parallelism": f"doc-shard x{world}",
                   "l2_flush": "inputs larger than L2 (matrix >> 126 MB)"},
        "e2e": {"value": e2e, "unit": "queries/s", "ms_per_step": wall_ms / K, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "stage_ms_per_step": {"scan": scan_ms / K, "scan_sweep_kernel": sweep_ms / K, "bm25": bm_ms / K, "fuse": fuse_ms / K, "comm": comm_ms / K},
    }
    # roofline of the dominant kernel
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload)
    if w["dim"] and scan_ms >= bm_ms:
        # dominant kernel = the sweep launch(es): CUDA events around those launches on the library's stream
        # (scan stage = threshold pass + sweep; its fraction is reported as batch_level_frac)
        ach = (scan_bytes / 1e9) / (max(sweep_ms, 1e-9) * 1e-3)
        n_local = hi - lo
        tflops = (2.0 * B * n_local * w["dim"] / 1e12) / (max(sweep_ms, 1e-9) / K * 1e-3) if tensor_core else None
        kname, kdesc = SCAN_VARIANTS.get(variant, ("emb_scan_kernel", "exact fp32 sweep"))
        line["scan"] = {"kernel": f"{kname} ({kdesc})",
                        "unproven_queries_rerun_exact_per_step": unproven / K,
                        "tensor_tflops_per_gpu": tflops}
        pk_json = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        if tensor_core and w.get("dtype") == "bf16" and B >= 512:
            tpeak = float(pk_json.get("bf16_tflops_sustained", 1400.0))
            line["roofline_tensor"] = {"kernel": kname, "bound": "tensor", "achieved": tflops, "peak": tpeak,
                                       "unit": "TFLOP/s", "frac": tflops / tpeak,
                                       "peak_source": "of measured (sustained)" if pk_json else "of fallback"}
        line["roofline"] = {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": peak,
                            "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "peak_source": f"of {peak_src}",
                            "kernel_ms_per_launch": sweep_ms / max(scan_launches, 1), "launches_per_step": scan_launches / K,
                            "algorithmic_bytes_per_launch": scan_bytes / max(scan_launches, 1),
                            "batch_level_frac": (scan_bytes / max(scan_launches, 1) * K / 1e9) / (scan_ms * 1e-3) / peak}
    else:
        ach = (postings * 8 / 1e9) / (bm_ms * 1e-3)
        line["roofline"] = {"kernel": "bm25_tile_kernel", "bound": "hbm", "achieved": ach, "peak": peak,
                            "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "peak_source": f"of {peak_src}",
                            "postings_per_s": postings / (bm_ms * 1e-3)}

    # ---- parity / recall of the timed configuration + CPU baseline (outside the timed region)
    if wl.get("chunked"):
        hits_planted = sum(int(h.doc_ids[0]) == int(pj) for h, pj in zip(hits, wl["planted"])) if rank == 0 else 0
        line["parity"] = {"planted_neighbour_is_rank1": hits_planted, "queries": B,
                          "note": "corpus generated chunk-wise (41 GB in fp32): no host copy for the CPU oracle; parity of this path is covered by tests/test_gpu_gemm.py::test_bf16_store_parity"}
    elif not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as orc
        orc.build()
        if world == 1:
            ix = orc.StrIndex(wl["data_all"]) if w["vocab"] else None
            st = orc.EmbStore(wl["rows"]) if w["dim"] else None
            cores = os.cpu_count() or 1
            done, tcpu, agree, tot = 0, 0.0, 0, 0
            recall_hits = recall_tot = 0
            while done < B and tcpu < args.cpu_seconds:
                m = min(cores, B - done)
                sb = orc.SearchBatch(ix, st)
                for i in range(done, done + m):
                    sb.add(mode, limit=10, similarity=0.0, q_vec=qv[i] if w["dim"] else None,
                           text=texts[i] if w["vocab"] else None)
                t1 = time.perf_counter()
                od, os_, on, oc = sb.run(cores)
                tcpu += time.perf_counter() - t1
                for k in range(m):
                    h = hits[done + k]
                    exp = set(od[k, :on[k]].tolist())
                    tot += 1
                    agree += (set(h.doc_ids.tolist()) == exp and h.count == int(oc[k])
                              and np.allclose(h.scores, os_[k, :on[k]], atol=1e-5, rtol=0))
                    recall_tot += len(exp)
                    recall_hits += len(exp & set(h.doc_ids.tolist()))
                done += m
            line["cpu_baseline"] = {"value": done / tcpu, "unit": "queries/s", "cores": cores, "kind": "port",
                                    "sample": f"first {done} of the {B} timed queries, one query per thread, "
                                              f"{tcpu:.1f} s of CPU work (C restatement of the reference algorithm)"}
            line["parity"] = {"queries_checked": tot, "identical_to_oracle": int(agree),
                              "recall_at_10_vs_oracle": recall_hits / max(recall_tot, 1)}
            if w["dim"]:
                rh = rt = 0
                for i in range(min(8, B)):
                    ed, ec = orc.vector_f64(st, qv[i], 10)
                    d, s, c = emb.search_batch(qv[i:i + 1], 10, -1.0)
                    got = set(d[0, :c[0]].tolist())
                    rt += len(ed)
                    rh += sum((int(x) in got) or abs(cc - ec[-1]) <= 1e-6 for x, cc in zip(ed, ec))
                line["parity"]["vector_recall_at_10_vs_fp64"] = rh / max(rt, 1)
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
