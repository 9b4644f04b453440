#!/usr/bin/env python
"""bench.py — milli query-time scoring path on B200 (contract: see the task's ④).

One *step* = one batch of 1024 typo-tolerant multi-term queries (SURVEY.md §8(d) cfg 2: synthetic "hackernews"
corpus, 1 M docs, one searchable field, default criteria, limit 20) through b200_search_batch.

  value     queries/sec computed from the CUDA-event time of the device work of the K steps
            (first kernel .. last kernel of every host<->device round trip; postings resident in HBM)
  e2e       queries/sec through the C ABI with HOST buffers: wall clock around the K calls, which contains the host-side
            ranking-rule control flow, every H2D/D2H copy and all synchronisation
  roofline  dominant kernel (largest accumulated CUDA-event time): algorithmic bytes / its event time vs measured HBM peak
  cpu_baseline  the CPU oracle ("port": C++ restatement of milli, the Rust reference cannot be built here) on a bounded
            sample of the same queries with all host threads

`--impl reference` times that CPU restatement alone, on the same workload/metric.
Multi-GPU (torchrun): the path shards by query — every rank holds a replica of the index and serves its own batch;
no data-path collective; value = all ranks' queries / max-over-ranks time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def ncu_traffic(kernel):
    """dram bytes per launch of `kernel` from the committed ncu capture (profiles/*_traffic.json, see tools/profile3.sh), or None"""
    names = {"eval_paths": "eval_dp_kernel", "scatter": "scatter_kernel", "lev_match": "lev_match_kernel", "act_compact": "act_compact_kernel",
             "pair_probe": "pair_probe_kernel", "emit": "emit_kernel"}
    try:
        import glob
        f = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")))[-1]
        return float(json.load(open(f))["kernels"][names[kernel]]["dram_bytes_per_launch"])
    except Exception:
        return None


def measured_peak_tflops():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["bf16_tflops"]), "measured (cuBLAS bf16 burst)"
        except Exception:
            pass
    return 1700.0, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(args, rank):
    from corpus.pyindexgen import IndexImage

    t = time.time()
    img = IndexImage(1)
    img.add_synthetic(args.docs, args.vocab, seed=0xB200)
    img.build()
    log(f"[rank {rank}] corpus: {img.n_docs} docs, {img.n_words} words, built in {time.time() - t:.1f}s")
    from meilisearch_b200.tokenizer import TokenBatch

    batches = [TokenBatch(img.synthetic_queries(args.batch, seed=1000 * rank + i)) for i in range(args.distinct_batches)]
    return img, batches


def run_reference(args, rank, world):
    """CPU arm: the oracle restatement of milli on all host threads, a bounded sample per step."""
    if rank != 0:
        return
    from oracle.pyoracle import OracleIndex

    img, batches = build_workload(args, 0)
    ix = OracleIndex(img)
    sample = min(args.batch, args.cpu_sample)
    from meilisearch_b200.tokenizer import TokenBatch

    qs = [TokenBatch(img.synthetic_queries(sample, seed=i)) for i in range(args.distinct_batches)]
    threads = best_cpu_run(ix, qs[0], sample)["cores"]
    for w in range(args.warmup):
        ix.search_batch(qs[w % len(qs)], n_threads=threads)
    t0 = time.perf_counter()
    lat = []
    for k in range(args.steps):
        r = ix.search_batch(qs[(args.warmup + k) % len(qs)], n_threads=threads)
        lat.append(r.seconds)
    dt = time.perf_counter() - t0
    qps = sample * args.steps / dt
    out = {
        "impl": "reference", "metric": "queries/sec (batch=1024, typo-tolerant multi-term keyword search, top-20)", "value": qps, "unit": "queries/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, img),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": f"{sample} queries per step x {args.steps} steps, {threads} threads, p50 {1e3 * float(np.median(np.concatenate(lat))):.2f} ms/query"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


def best_cpu_run(oracle, tokens, sample):
    """One query per thread (the reference's own concurrency model); the thread count that gives the best QPS is reported."""
    cores = os.cpu_count() or 1
    best = None
    for threads in sorted({cores, max(1, cores // 2), max(1, cores // 4)}, reverse=True):
        tc = time.perf_counter()
        r = oracle.search_batch(tokens, n_threads=threads)
        dt = time.perf_counter() - tc
        cand = {"value": sample / dt, "unit": "queries/s", "cores": threads, "kind": "port",
                "sample": f"{sample} queries of the timed workload, {threads} of {cores} host threads (best of {cores}, {cores // 2}, {cores // 4}), "
                          f"p50 {1e3 * float(np.median(r.seconds)):.2f} ms/query; CPU restatement of milli, not milli itself"}
        if best is None or cand["value"] > best["value"]:
            best = cand
    return best


def workload_config(args, img):
    return {"workload": f"cfg2 hackernews-like synthetic: {img.n_docs} docs x 1 field, {img.n_words}-word dictionary, batch={args.batch} queries of 2-4 words "
                        "(40% clean / 40% one edit / 20% two edits, last word prefix p=0.3), criteria words,typo,proximity,attributeRank,wordPosition,exactness, "
                        "TermsMatchingStrategy::Last, limit 20",
            "batch": args.batch, "docs": int(img.n_docs), "l2": "working set (posting store + per-batch matrices) exceeds the 126 MB L2; distinct query batch every step"}


def sharded_vector_stage(ix, rank, world, local_rank):
    """cfg 5 shape, weak scaling: every rank owns 1e6 x 768 fp16 rows of a (world x 1e6)-row matrix (contiguous docid ranges), scans
    them for the SAME 1024 queries (tcgen05 GEMM + fused top-100), then one NCCL all-gather of the per-shard top-100 and a merge."""
    import torch
    import torch.distributed as dist

    from meilisearch_b200.parallel import merge_sharded_topk

    try:
        n, dim, B, k = 1_000_000, 768, 1024, 100
        rng = np.random.default_rng(0xE5BED + rank)
        ix.set_embeddings(rng.standard_normal((n, dim), dtype=np.float32), np.arange(rank * n, (rank + 1) * n, dtype=np.uint32))
        q = np.random.default_rng(7).standard_normal((B, dim), dtype=np.float32)
        dev = torch.device("cuda", local_rank)

        def step():
            ids, dst, cnt = ix.nns_by_vector(q, k)
            return merge_sharded_topk(ids.astype(np.int64), dst, cnt.astype(np.int64), k, device=dev)

        for _ in range(3):
            step()
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            m_ids, m_dst, m_cnt = step()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return {"workload": f"corpus-sharded: {world} x (1e6 x 768 fp16) rows by docid range, 1024 queries, top-100; one all-gather of "
                            f"{world} x 1024 x 100 x (i64 docid, f32 distance) + merge", "rows_total": n * world,
                "ms_per_batch": 1e3 * float(dt[0]) / reps, "queries_per_s": B * reps / float(dt[0]),
                "all_sorted": bool((m_dst[:, 1:] >= m_dst[:, :-1]).all().item())}
    except Exception as e:  # secondary measurement
        return {"error": str(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--docs", type=int, default=1_000_000)
    ap.add_argument("--vocab", type=int, default=400_000)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--distinct-batches", type=int, default=4)
    ap.add_argument("--cpu-sample", type=int, default=512)
    ap.add_argument("--no-vector", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import meilisearch_b200 as mb

    mb.load_library()  # fails loudly if the CUDA extension is missing
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ranks share the host: keep the per-rank worker pools inside the machine's cores
    os.environ.setdefault("B200_HOST_THREADS", str(max(8, min(32, (os.cpu_count() or 64) // max(1, world)))))
    img, batches = build_workload(args, rank)
    t = time.time()
    ix = mb.Index(img, device=local_rank)
    log(f"[rank {rank}] staged {ix.stats()['hbm_bytes_staged'] / 1e6:.0f} MB to HBM in {time.time() - t:.1f}s")

    def step(i):
        res = ix.search().query(batches[i % len(batches)]).execute()
        return res

    for w in range(args.warmup):
        res = step(w)
    # parity on a sample of the first batch against the oracle (checker only; not in the timed region)
    parity = None
    if rank == 0:
        from meilisearch_b200.tokenizer import TokenBatch
        from oracle.pyoracle import OracleIndex

        sample_q = img.synthetic_queries(args.batch, seed=1000 * rank + 0)[: min(128, args.batch)]
        tb = TokenBatch(sample_q)
        got = ix.search().query(tb).execute()
        want = OracleIndex(img).search_batch(tb, n_threads=os.cpu_count() or 1)
        mism = sum(1 for q in range(len(sample_q)) if got.ids(q) != want.ids(q))
        parity = {"checked": len(sample_q), "top20_mismatches": mism}
        log(f"parity: {parity}")

    sampler = ClockSampler(local_rank)
    os.environ["B200_KERNEL_TIMERS"] = "0"  # the e2e region runs as production would: no per-kernel event records
    ix.reset_stats()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    t0 = time.perf_counter()
    lat = []
    for k in range(args.steps):
        ts = time.perf_counter()
        res = step(args.warmup + k)
        lat.append(time.perf_counter() - ts)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    st_e2e = ix.stats()
    n_ok = int((res.status == 0).sum())
    # second timed region, software pipeline off (one lane): kernels of different lanes no longer overlap, so the CUDA-event
    # intervals are clean.  `value` and the roofline come from this pass; `e2e` from the pipelined pass above.
    os.environ["B200_SINGLE_LANE"] = "1"
    os.environ["B200_KERNEL_TIMERS"] = "1"
    step(args.warmup)
    ix.reset_stats()
    torch.cuda.synchronize()
    for k in range(args.steps):
        step(args.warmup + k)
    torch.cuda.synchronize()
    del os.environ["B200_SINGLE_LANE"]
    st = ix.stats()
    dev_s = st["device_ms"] / 1e3 + sum(st["kernels"][k]["ms"] for k in ("lev_match",)) / 1e3
    if world > 1:
        tt = torch.tensor([wall, dev_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        wall, dev_s = float(tt[0]), float(tt[1])
    total_q = args.batch * args.steps * world
    sharded = None
    if world > 1 and not args.no_vector:
        sharded = sharded_vector_stage(ix, rank, world, local_rank)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel
    peak, peak_kind = measured_peak_gbs()
    kern = {k: v for k, v in st["kernels"].items() if v["count"]}
    dom = max(kern, key=lambda k: kern[k]["ms"])
    d = kern[dom]
    per_launch_ms = d["ms"] / d["count"]
    achieved = (d["bytes"] / d["count"]) / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "peak_source": peak_kind, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(dom), "launches": int(d["count"]), "avg_launch_ms": per_launch_ms,
                "algorithmic_bytes_per_launch": d["bytes"] / d["count"],
                "kernel_time_share": {k: round(v["ms"] / max(1e-9, sum(x["ms"] for x in kern.values())), 4) for k, v in kern.items()}}

    # CPU baseline: oracle on a bounded sample of the same workload, all host threads
    from meilisearch_b200.tokenizer import TokenBatch
    from oracle.pyoracle import OracleIndex

    sample = min(args.batch, args.cpu_sample)
    sq = TokenBatch(img.synthetic_queries(args.batch, seed=1000 * rank + (args.warmup % len(batches)))[:sample])
    o = OracleIndex(img)
    cpu = best_cpu_run(o, sq, sample)

    out = {
        "metric": "queries/sec (batch=1024, typo-tolerant multi-term keyword search, top-20)", "value": total_q / dev_s, "unit": "queries/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, img),
        "e2e": {"value": total_q / wall, "unit": "queries/s", "ms_per_step": 1e3 * wall / args.steps, "p50_batch_ms": 1e3 * float(np.median(lat)),
                "h2d_bytes_per_step": int(st_e2e["h2d_bytes"] / args.steps), "d2h_bytes_per_step": int(st_e2e["d2h_bytes"] / args.steps),
                "device_steps_per_batch": st_e2e["device_steps"] / args.steps, "lanes": os.environ.get("B200_DRIVERS", "4") + "x" + os.environ.get("B200_LANES_PER_DRIVER", "1")},
        "gpu_launches": int(st_e2e["kernel_launches"]),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "parity": parity,
        "queries_ok": n_ok,
        "host_ms_per_step": {k: v / args.steps for k, v in st_e2e["host_ms"].items()},
        "algorithmic_bytes_per_step": {"posting": int(st["posting_bytes"] / args.steps), "matrix": int(st["matrix_bytes"] / args.steps),
                                       "dictionary": int(st["dictionary_bytes"] / args.steps)},
    }

    # secondary: the vector stage (cfg 4, d=768 fp16, top-100) — GEMV roofline
    if not args.no_vector and world == 1:
        try:
            rng = np.random.default_rng(0xE5BED)
            n, dim = 1_000_000, 768
            emb = rng.standard_normal((n, dim), dtype=np.float32)
            ix.set_embeddings(emb)
            del emb
            q = rng.standard_normal((8, dim), dtype=np.float32)
            for _ in range(3):
                ix.nns_by_vector(q[:1], 100)
            ix.reset_stats()
            tv = time.perf_counter()
            reps = 20
            for i in range(reps):
                ix.nns_by_vector(q[i % 8: i % 8 + 1], 100)
            wall_v = time.perf_counter() - tv
            sv = ix.stats()["kernels"]["vec_dist"]
            gbs = sv["bytes"] / (sv["ms"] * 1e-3) / 1e9
            out["vector_stage"] = {"workload": "cfg4: 1e6 x 768 fp16 rows, B=1 cosine top-100 (matrix 1.5 GB > L2)", "kernel": "vec_dist",
                                   "avg_launch_ms": sv["ms"] / sv["count"], "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak,
                                   "e2e_queries_per_s": reps / wall_v}
            # batched vector stage (B=1024): tcgen05 GEMM with the top-100 fused into its epilogue, end to end from host buffers
            qb = rng.standard_normal((1024, dim), dtype=np.float32)
            for _ in range(3):
                ix.nns_by_vector(qb, 100)
            ix.reset_stats()
            tv = time.perf_counter()
            reps = 10
            for i in range(reps):
                ix.nns_by_vector(qb, 100)
            wall_b = time.perf_counter() - tv
            sg = ix.stats()["kernels"]["vec_gemm_topk"]
            if sg["count"]:
                ms = sg["ms"] / sg["count"]
                tflops = 2.0 * 1024 * n * dim / (ms * 1e-3) / 1e12
                tpeak = measured_peak_tflops()
                out["vector_stage_batched"] = {"workload": "cfg4 batched: 1024 queries x (1e6 x 768 fp16), cosine top-100, fp16 operands / fp32 accumulate",
                                               "kernel": "vec_gemm_topk (+vec_merge)", "avg_launch_ms": ms,
                                               "roofline": {"bound": "tensor", "achieved": tflops, "peak": tpeak[0], "peak_source": tpeak[1],
                                                            "unit": "TFLOP/s", "frac": tflops / tpeak[0]},
                                               "kernel_queries_per_s": 1024 / (ms * 1e-3), "e2e_queries_per_s": 1024 * reps / wall_b}
            # hybrid (execute_hybrid, semanticRatio 0.5): the timed keyword batches + one query vector each, end to end
            hv = rng.standard_normal((args.batch, dim), dtype=np.float32)
            for w in range(2):
                ix.search().query(batches[w % len(batches)]).semantic(hv).execute_hybrid(0.5)
            th = time.perf_counter()
            reps = 4
            lat_h = []
            for i in range(reps):
                t1 = time.perf_counter()
                rh = ix.search().query(batches[i % len(batches)]).semantic(hv).execute_hybrid(0.5)
                lat_h.append(time.perf_counter() - t1)
            wall_h = time.perf_counter() - th
            out["hybrid_stage"] = {"workload": "cfg2 keyword batch + 1 query vector per query over 1e6 x 768 fp16 embeddings, semanticRatio 0.5, limit 20",
                                   "e2e_queries_per_s": args.batch * reps / wall_h, "p50_batch_ms": 1e3 * float(np.median(lat_h)),
                                   "queries_ok": int((rh.status == 0).sum())}
        except Exception as e:  # the headline number must not die with the secondary one
            out["vector_stage"] = {"error": str(e)}
    if sharded is not None:
        out["vector_stage_sharded"] = sharded
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
