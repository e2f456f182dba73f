#!/usr/bin/env python
"""bench.py -- the Myrio hot path on N B200s (BASELINE.json: query x reference-leaf
similarity scores/s and k-mers counted/s at 1/2/4/8 B200).

  python bench.py --gpus N --steps K --warmup W            # the CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference

Headline workload (BASELINE.json configs[1]): a synthetic BOLD-scale ITS tree of 100 000
leaves per GPU x 100 000 simulated reads, k = 18.  One STEP = one pass of the hot path over
the read batch: 2-bit k-mer counting of every read (K1), L2 normalisation, scoring of every
read against every leaf of the rank's shard over the inverted index (K4), top-100 per read
(K5) and -- for N > 1 -- the bound all-reduce, the all-gather of the shard lists and the merge
(K7), all inside libmyrio_b200.so (myrio_score_topx_sharded).  The reference tree is counted
(K2) and indexed (K3) once before the timed region; both are timed separately and reported.

Data are independent of N: rank r holds sub-tree r of one global tree (same root, payload ids
r*L .. (r+1)*L), and the ONE read batch is simulated from sub-tree 0, so N = 1 is exactly
configs[1] and larger N add sister clades to the reference ("weak": leaves per GPU fixed).
`strong` (BASELINE.json configs[4]) is the 1 M-leaf tree -- 8 fixed sub-trees of 125 000
leaves -- sharded over the N ranks against one fixed read batch drawn from all of it.

`value` has the reads resident in HBM; `e2e` goes through the public API with pinned host
buffers (H2D of the packed reads on every rank, D2H of each rank's 1/N of the merged top-x
inside the timed region).  Inputs (index >= 1 GB per GPU) are larger than the 126 MB L2.

stdout carries exactly one JSON line; everything else any library prints (NCCL_DEBUG output
included -- the caller's NCCL_DEBUG setting is left alone) goes to stderr.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "query_x_leaf_similarity_scores_per_s"
UNIT = "scores/s"
N_STRONG_SUBTREES = 8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--leaves", type=int, default=100_000, help="reference leaves per GPU (weak scaling)")
    ap.add_argument("--reads", type=int, default=100_000, help="reads (queries), total")
    ap.add_argument("--k", type=int, default=18)
    ap.add_argument("--x", type=int, default=100)
    ap.add_argument("--tile-leaves", type=int, default=0)
    ap.add_argument("--topx-batch", type=int, default=0, help="N > 1: cap the queries per exchange batch (tests)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clustering", action="store_true", help="skip the clustering stage (K6) measurement")
    ap.add_argument("--clusters", type=int, default=4)
    ap.add_argument("--cluster-reps", type=int, default=3, help="timed cluster() calls per variant (median reported)")
    ap.add_argument("--cluster-warmup", type=int, default=1)
    ap.add_argument("--strong-leaves", type=int, default=1_000_000,
                    help="total leaves of the strong-scaling tree (configs[4]); 0 skips the block")
    ap.add_argument("--strong-reads", type=int, default=100_000)
    ap.add_argument("--strong-steps", type=int, default=2)
    ap.add_argument("--no-variants", action="store_true",
                    help="skip the sparse-path variants (tree without conserved block, 1 %% ambiguity codes, k=12)")
    ap.add_argument("--variant-reads", type=int, default=50_000)
    return ap.parse_args()


def config_dict(args, world):
    return {
        "workload": f"configs[1]: synthetic BOLD-Plantae-scale ITS tree, {args.leaves} leaves per GPU x "
                    f"{args.reads} reads, k={args.k}, top-{args.x}, cosine",
        "leaves_per_gpu": args.leaves, "leaves_total": args.leaves * world, "reads": args.reads,
        "k": args.k, "x": args.x, "read_cutoff_t1": 0.1, "nb_resamples": 32,
        "topx": f"fused top-{args.x} (the cut() of tax/results.rs:310-329): every (query, leaf) pair's shared k-mers are "
                "counted and its score bounded from above; a score is finished to its exact f32 value only if the bound "
                "reaches the query's running x-th best -- the returned top-x is bit-identical to a full scoring; "
                "`all_scores_finished` is the same step with every score finished",
        "parallelism": f"leaf-shard x{world} (contiguous payload-id blocks), queries replicated; seed pass, "
                       f"ncclAllReduce(MAX) of the per-query bounds, sweep, ncclAllGather of 8-byte composites, merge "
                       f"-- all inside myrio_score_topx_sharded" if world > 1 else "single GPU",
        "data_independent_of_n": "rank r indexes sub-tree r (seed 1002, shard r) of one global tree; the read batch "
                                 "(seed 2002) is simulated from sub-tree 0 at every N",
        "l2_policy": "inputs larger than L2 (index >= 1 GB per GPU vs 126 MB L2); no explicit flush",
        "seeds": {"tree": 1002, "reads": 2002},
    }


# ------------------------------------------------------------------ stdout discipline

_REAL_STDOUT = None


def claim_stdout():
    """fd 1 -> stderr for the rest of the process (NCCL's C-level logging follows), the JSON line goes to
    the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


# ------------------------------------------------------------------ clocks

class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md's clocks line).  Read in-process
    through NVML (the library nvidia-smi itself reads) every 500 ms by a thread that holds no lock of ours:
    an `nvidia-smi -lms 200` child next to several busy ranks stalled CUDA calls for 100-800 ms a few times per
    second on these boxes (profiles/r02_smi_hiccups.md), which is a property of the probe, not of the step.
    Falls back to the nvidia-smi child (at -lms 1000) when pynvml is unavailable."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None
        self.sm, self.reasons, self.mx, self.stop_flag, self.thread, self.source = [], set(), None, False, None, None
        self.period = float(os.environ.get("MYRIO_BENCH_NVML_PERIOD", "0.5"))
        self.device_mhz = []  # on-device measurements inside the timed regions (myrio_bench_clock_probe)

    def _nvml_loop(self, nv, h):
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        mode = os.environ.get("MYRIO_BENCH_NVML", "full")  # diagnosis: "clock" skips the event-reason query
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                if mode != "clock":
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                    for name, bit in names:
                        if mask & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def start_inline(self):
        """In-line mode (default): no thread; `poll()` (NVML clock + throttle reasons) is called by the timing loop
        at the two ends of every timed region, the SM clock inside the region comes from on-device probes.  A
        concurrent NVML query from another thread or process stalls this process's CUDA calls for 20-900 ms every
        now and then (profiles/r02_smi_hiccups.md); the duration of every poll is recorded (`poll_ms`)."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else self.gpu
            self._nv, self._h = nv, nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM))
            self.source = ("sm_mhz: on-device probes (SM cycles per %globaltimer ns over ~0.5 ms) enqueued between steps INSIDE "
                           "the timed regions; sm_max_mhz / reasons: NVML in-process (pynvml) before the first timed region "
                           "and right after the last event of each")
            self.inline, self.poll_ms = True, []
        except Exception:
            self.inline = False
            self.start()

    def poll(self):
        if not getattr(self, "inline", False):
            return
        nv = self._nv
        t0 = time.perf_counter()
        try:
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
            for name, bit in (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                              ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                              ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                              ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)):
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass
        self.poll_ms.append((time.perf_counter() - t0) * 1e3)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # LOCAL_RANK indexes CUDA_VISIBLE_DEVICES; NVML wants the physical index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else self.gpu
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.source = "NVML in-process (pynvml), every 500 ms"
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.source = "nvidia-smi -lms 1000"
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "1000",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if getattr(self, "inline", False):
            if not self.sm and not self.device_mhz:
                return {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": [], "samples": 0, "source": self.source}
            under_load = self.device_mhz or self.sm
            return {"sm_mhz": float(np.median(under_load)), "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                    "samples": len(under_load), "source": self.source,
                    "sm_mhz_device_probes": [round(m, 1) for m in self.device_mhz],
                    "sm_mhz_nvml_at_region_boundaries": self.sm,
                    "poll_ms_max": round(max(self.poll_ms), 3) if self.poll_ms else None,
                    "poll_ms_total": round(sum(self.poll_ms), 3) if self.poll_ms else None}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if not self.sm:
                return {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": [], "samples": 0, "source": self.source}
            return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                    "samples": len(self.sm), "source": self.source}
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                   r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "source": self.source}


# ------------------------------------------------------------- data

def weak_tree(args, shard):
    from myrio_b200 import synth
    return synth.make_tree(args.leaves, 1002, root_seed=1002, shard=shard)


def weak_reads(args, tree0):
    from myrio_b200 import synth
    return synth.make_reads(tree0, args.reads, 2002)


def strong_subtree(args, j):
    """Sub-tree j of the configs[4] tree: fixed for every N (the tree is ALWAYS cut into 8 sub-trees, a rank
    holds 8/N consecutive ones)."""
    from myrio_b200 import synth
    per = args.strong_leaves // N_STRONG_SUBTREES
    return synth.make_tree(per, 1005, root_seed=1005, shard=j)


def strong_reads_of(args, sub, j):
    from myrio_b200 import synth
    return synth.make_reads(sub, args.strong_reads // N_STRONG_SUBTREES, 2005 + j)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def latest_profile(pattern: str):
    """Newest committed profiles/*<pattern>*.json (files are named per round: r01_, r02_, ...)."""
    d = os.path.join(ROOT, "profiles")
    best = None
    if os.path.isdir(d):
        for f in sorted(os.listdir(d)):
            if f.endswith(".json") and pattern in f:
                try:
                    best = (f, json.load(open(os.path.join(d, f))))
                except Exception:
                    pass
    return best


def ascii_leaves(tree):
    """FASTA sequence text of an ACGT tree (what load_from_fasta_string reads), for the device ingestion."""
    from myrio_b200 import synth
    return synth.ASCII[tree.codes & 3], np.ascontiguousarray(tree.base_off, np.uint64)


def logical_leaf_kmers(tree, k):
    return float(2 * (tree.lengths() - k + 1).clip(0).sum())


def k1_algorithmic_bytes(reads, nnz):
    """ceil(n/4) + n in, 12*D + 16 out per read."""
    n = reads.lengths()
    return float(((n + 3) // 4).sum() + n.sum() + 12 * nnz + 16 * reads.n)


def k2_algorithmic_bytes(tree, nnz):
    """ceil(n/2) in, 10*D + 12 out per leaf."""
    n = tree.lengths()
    return float(((n + 1) // 2).sum() + 10 * nnz + 12 * tree.n)


# ------------------------------------------------------------- clustering stage (K6)

SMEM_BYTES_PER_CLK_PER_SM = 128.0  # nominal shared-memory bandwidth of one SM (32 banks x 4 B)


def bench_clustering(ctx, api, resident, n_clusters, sm_mhz, cpu_reads=None, cpu_seconds=0.0, world=1,
                     cluster_fn=None, reduce_max=None, reduce_sum=None, reps=3, warmup=1):
    """Stage 3 of the hot path: clustering::cluster (clustering.rs:55-270) on the whole read batch at
    the reference defaults (k=6, t1=0.2, t2=20), first without and then with silhouette trimming
    (ssdcf 1.4, config_default.toml:30) -- the O(N^2) reads x reads similarity.  Kernel times are CUDA
    events on the library's stream; the pair / term counts are exact (myrio_cluster_stats)."""
    cluster_fn = cluster_fn or api.cluster
    reduce_max = reduce_max or (lambda v: v)
    reduce_sum = reduce_sum or (lambda v: v)
    out = {"workload": f"cluster() k=6 t1=0.2 t2=20 on {resident.n} reads (the same batch at every N), {n_clusters} "
                       f"clusters (farthest-point fill), cosine; then the same with silhouette trimming 1.4",
           "parallelism": "single GPU" if world == 1 else
                          f"reads replicated, rows of every similarity tile split over {world} GPUs, per-row results "
                          f"exchanged with ncclAllGather inside the library (myrio_cluster_sharded, no callback)",
           "scaling": "strong"}
    # measured shared-memory bandwidth (a conflict-free LDS.128 read kernel of the library), the K6 denominator
    smem_meas = ctx.bench_smem_bandwidth()
    for name, sil in (("assign_only", None), ("with_silhouette", 1.4)):
        prm = api.ClusteringParameters(k=6, t1=0.2, t2=20, expected_nb_of_clusters=n_clusters,
                                       silhouette_trimming=sil)
        for _ in range(warmup):
            cluster_fn(resident, prm)  # warm-up (allocator pool, kernel attributes)
        calls = []
        for _ in range(max(1, reps)):  # median of the timed calls (the stream-ordered allocator adds run-to-run noise)
            ctx.prof_reset()
            ctx.prof_enable(True)
            reduce_max(0.0)  # barrier
            t0 = time.perf_counter()
            res = cluster_fn(resident, prm)
            calls.append(reduce_max(time.perf_counter() - t0))  # max over ranks
        call_s = sorted(calls)[len(calls) // 2]
        st = ctx.cluster_stats()  # this rank's share of the tiles
        pairs_all = int(reduce_sum(float(st["pair_similarities"])))
        tile_ms = sum(ctx.prof_query(p)[0] for p in ("k6_pair_", "k6_silsum", "k6_argmax"))
        sil_ms = ctx.prof_query("k6_pair_silhouette")[0]
        all_ms, n_launch = ctx.prof_query("")
        ctx.prof_enable(False)
        ctx.prof_reset()
        d = {"call_ms": call_s * 1e3, "calls_ms": [c * 1e3 for c in calls], "kernels_ms": all_ms, "k6_tile_kernels_ms": tile_ms, "gpu_launches": n_launch,
             "iters": res.n_iters, "cluster_sizes": [int(len(c)) for c in res.clusters],
             "rejected": int(len(res.rejected)), "pair_similarities": pairs_all,
             "pair_similarities_rank0": st["pair_similarities"],
             "terms_rank0": st["terms"], "pair_similarities_per_s": pairs_all / max(call_s, 1e-9)}
        if getattr(res, "exchange", None):
            d["exchange"] = res.exchange
        if sil is not None and sil_ms > 0:
            # dominant kernel of the stage: pair_tile_canon_kernel.  Every multiply-add term reads one f32 of a
            # column from shared memory (4 B); the bound is the SMs' shared-memory bandwidth.
            nominal = 148 * SMEM_BYTES_PER_CLK_PER_SM * (sm_mhz or 1965.0) * 1e6 / 1e9
            peak = smem_meas["gbps"] if smem_meas["gbps"] > 0 else nominal
            ach = 4.0 * st["terms"] / (sil_ms * 1e-3) / 1e9
            d["roofline"] = {"bound": "shared-memory bandwidth", "kernel": "k6_pair_silhouette (pair_tile_canon_kernel)",
                             "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "kernel_ms": sil_ms, "terms_per_s": st["terms"] / (sil_ms * 1e-3),
                             "peak_source": "measured in this run: myrio_bench_smem_bandwidth (conflict-free LDS.128 reads "
                                            f"on every SM, {smem_meas['ms']:.2f} ms)",
                             "peak_nominal": nominal,
                             "note": "algorithmic bytes = 4 B (one column value read from shared memory) per multiply-add "
                                     "term; the scheduled slab is conflict-free, the loss is its ~1.12x waiting slots, "
                                     "column staging and the row-slab stream (profiles/r01_k6_canon_ncu_summary.json)"}
        out[name] = d
    if cpu_reads is not None and cpu_seconds > 0:
        import oracle
        threads = oracle.max_threads()
        n = 256
        oc = oracle.count_reads(cpu_reads.codes, cpu_reads.qual, cpu_reads.base_off, 6, 0.2)
        on = oracle.normalize(oc, threads)
        rows = oracle.Csr(on.row_off[:n + 1], on.keys[:int(on.row_off[n])], on.vals[:int(on.row_off[n])])
        t0 = time.perf_counter()
        oracle.pairwise(rows, on, threads=threads)
        dt = max(time.perf_counter() - t0, 1e-9)
        reps = int(max(1, min(20, cpu_seconds / dt)))
        t0 = time.perf_counter()
        for _ in range(reps):
            oracle.pairwise(rows, on, threads=threads)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": reps * n * on.n_rows / dt, "unit": "pair similarities/s", "cores": threads,
                               "kind": "port", "sample": f"{reps} x ({n} x {on.n_rows}) read x read merge-join dots at k=6, "
                                                         f"{threads} OpenMP threads"}
    return out


# ------------------------------------------------------------- CPU arms

def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port) on ALL host cores, rank 0 only; the same
    config (tree seed 1002 shard 0 = the per-GPU shard, reads seed 2002) at every N."""
    if rank != 0:
        return
    claim_stdout()
    import oracle
    from myrio_b200 import synth
    oracle.set_num_threads(host_threads())  # torchrun exports OMP_NUM_THREADS=1
    threads = oracle.max_threads()
    tree = weak_tree(args, 0)
    reads = weak_reads(args, tree)
    # leaf vectors (the reference's "index" is the list of normalised leaf vectors)
    t0 = time.perf_counter()
    lc = oracle.count_leaves(synth.tree_to_iupac(tree), tree.base_off, args.k, 32, threads=threads)
    t_leaves = time.perf_counter() - t0
    ln = oracle.normalize(lc, threads)
    # queries
    nq_probe = min(reads.n, 8)
    sub = reads.subset(np.arange(min(reads.n, 4096)))
    t0 = time.perf_counter()
    rc = oracle.count_reads(sub.codes, sub.qual, sub.base_off, args.k, 0.1, threads=1)  # serial like the reference
    t_reads = time.perf_counter() - t0
    t0 = time.perf_counter()
    oracle.count_reads(sub.codes, sub.qual, sub.base_off, args.k, 0.1, threads=threads)
    t_reads_par = time.perf_counter() - t0
    qn = oracle.normalize(rc, threads)

    def score_n(n):
        t0 = time.perf_counter()
        for i in range(n):
            k, v = qn.row(i)
            s = oracle.score_all(ln, k, v, threads=threads)
            oracle.topx(s, args.x)
        return time.perf_counter() - t0

    t_probe = score_n(nq_probe)
    per_q = t_probe / nq_probe
    n_step = int(max(4, min(qn.n_rows, args.cpu_seconds / max(per_q, 1e-9) / max(1, args.steps + args.warmup))))
    for _ in range(args.warmup):
        score_n(n_step)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        score_n(n_step)
    dt = time.perf_counter() - t0
    # the work a rank-count of N asks for is Q x (N x L) scores; the CPU arm scores against one L-leaf shard
    # per step sample, so its rate is per (query, leaf) pair either way
    value = n_step * args.steps * tree.n / dt
    sample = (f"{n_step} queries x {tree.n} leaves per step (first reads of the seed-2002 batch), "
              f"{threads} OpenMP threads over leaves like rayon (tax/results.rs:89-93)")
    leaf_kmers = logical_leaf_kmers(tree, args.k)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_dict(args, max(world, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "kmers_counted_per_s": {
            "reads_serial_like_the_reference": float(rc.aux.sum()) / t_reads,
            "reads_all_threads": float(rc.aux.sum()) / t_reads_par,
            "leaves_reference_equivalent_x32": leaf_kmers * 32 / t_leaves,
            "leaves_logical": leaf_kmers / t_leaves,
            "cores": threads, "sample": f"{sub.n} reads; all {tree.n} leaves x 32 resamples",
        },
        "note": "CPU restatement of the reference algorithm (the Rust binary cannot be built: no cargo/rustc)",
    }
    emit(line)


# ------------------------------------------------------------- GPU arm

def run_ours(args, rank, local_rank, world):
    claim_stdout()
    import torch
    import torch.distributed as dist
    from myrio_b200 import _lib, api, synth
    from myrio_b200 import dist as mdist

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = api.Context(local_rank)
    L = ctx._L
    dev = torch.device("cuda", local_rank)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    comm_info = mdist.init_library_comm(ctx) if world > 1 else ctx.comm_info()

    def barrier():
        if world > 1:
            dist.barrier()
        ctx.sync()
        torch.cuda.synchronize(dev)

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    step_trace = {}
    sampler = None  # ClockSampler of rank 0 while the headline regions are timed

    def timed(fn, steps, tag=None):
        """`steps` calls bracketed by barrier + synchronize, CUDA events on the library's stream, max over ranks.
        With a tag the end of every step is marked too (diagnostic: rank 0's per-step times in the JSON line)."""
        barrier()
        import gc
        gc.collect()
        gc.disable()  # no collector pauses on any rank while the K steps are timed (re-enabled below)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = []
        # Clocks.  INSIDE the timed region the SM clock is measured on the device (myrio_bench_clock_probe: SM
        # cycles against %globaltimer over ~0.5 ms, enqueued on the library's stream between two steps) -- no
        # driver call, so nothing is perturbed.  NVML (clock + throttle reasons) is polled once before the first
        # region (followed by one more warm-up step) and immediately after the last event of every region: a
        # query cost 2-100 ms on these boxes, slowed the step after it, and stalled other ranks for 100-900 ms
        # when it ran inside a region (profiles/r02_smi_hiccups.md).
        probe_at = []
        if tag and sampler is not None:  # N > 1: one probe, before the last step (rank 0 must not drift from its peers)
            probe_at = sorted({steps // 4, steps // 2, (3 * steps) // 4, steps - 1}) if world == 1 else [steps - 1]
        e0.record(ext)
        n_probe = 0
        for it in range(steps):
            if it in probe_at:
                ctx.clock_probe(n_probe)
                n_probe += 1
            fn()
            if tag:
                m = torch.cuda.Event(enable_timing=True)
                m.record(ext)
                marks.append(m)
        ctx.sync()
        torch.cuda.synchronize(dev)
        e1.record(ext)
        e1.synchronize()
        if tag and sampler is not None:
            sampler.poll()
            sampler.device_mhz.extend(m for m in ctx.clock_read(n_probe) if m > 0)
        gc.enable()
        ms = allmax(e0.elapsed_time(e1))
        if tag:
            ts = [e0.elapsed_time(m) for m in marks]
            step_trace[tag] = [round(b - a, 3) for a, b in zip([0.0] + ts[:-1], ts)]
        barrier()
        return ms

    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()

    tree0 = weak_tree(args, 0)
    tree = tree0 if rank == 0 else weak_tree(args, rank)
    reads = weak_reads(args, tree0)
    Q, x, k = reads.n, args.x, args.k

    # ---- setup (timed separately): FASTA text -> packed leaves on the device (K0), K2 leaf counting,
    # normalise, K3 index build
    ctx.prof_enable(True)
    text, text_off = ascii_leaves(tree)
    t0 = time.perf_counter()
    store, _ = api.TaxTreeStore.from_fasta_records(ctx, text, text_off, leaf_id_base=rank * args.leaves)
    t_upload_leaves = time.perf_counter() - t0
    store.compute_and_append_kmer_counts(k, 32)  # first call: kernel attributes, pool growth
    ctx.sync()
    store.kmer_store_counts_vec, store.k_precomputed = [], []
    ctx.prof_reset()
    t0 = time.perf_counter()
    store.compute_and_append_kmer_counts(k, 32)
    ctx.sync()
    t_count_leaves = time.perf_counter() - t0
    k2_ms, _ = ctx.prof_query("k2_")
    leaf_nnz = store.kmer_store_counts_vec[0].nnz
    t0 = time.perf_counter()
    ttc = api.TaxTreeCompute.from_store_tree(store, k, 32, tile_leaves=args.tile_leaves)
    ctx.sync()
    t_index = time.perf_counter() - t0
    info = ttc.info()
    leaf_kmers = logical_leaf_kmers(tree, k)
    ctx.prof_enable(False)
    ctx.prof_reset()

    # ---- pinned host copies of the packed reads (e2e path) and resident reads (value path)
    words, word_off = synth.pack_dna2(reads)
    base_off = np.ascontiguousarray(reads.base_off, np.uint64)
    keep = [pinned(words), pinned(word_off), pinned(base_off), pinned(reads.qual)]
    p_words, p_woff, p_boff, p_qual = (kp[1] for kp in keep)
    h2d_rank = int(p_words.nbytes + p_woff.nbytes + p_boff.nbytes + p_qual.nbytes)
    # every rank brings back its 1/N of the merged result: together the host holds all of it once
    q_lo, q_hi = mdist.shard_bounds(Q, rank, world)
    out_s_t = torch.empty((max(q_hi - q_lo, 1), x), dtype=torch.float32).pin_memory()
    out_i_t = torch.empty((max(q_hi - q_lo, 1), x), dtype=torch.int32).pin_memory()
    d2h_total = int(Q * x * 8)
    resident = api.MyrSeqs.from_packed(ctx, p_words, p_woff, p_boff, p_qual)

    # result tensors of the step, allocated once (a caller's loop owns its output buffers)
    res_s = torch.empty((Q, x), dtype=torch.float32, device=dev)
    res_i = torch.empty((Q, x), dtype=torch.int32, device=dev)
    torch.cuda.synchronize(dev)
    sync_mode = os.environ.get("MYRIO_BENCH_STEP_SYNC", "")  # diagnosis: "device" ends a resident step with a device-wide sync

    def topx(qn):
        if world > 1:
            return mdist.sharded_topx(ctx, ttc, qn, x, api.SIM_COSINE, max_batch=args.topx_batch, out=(res_s, res_i))
        return mdist.sharded_topx_allgather(ctx, ttc, qn, x, api.SIM_COSINE, out=(res_s, res_i))  # one rank: K4 + K5

    def step_resident():
        counts = resident.compute_kmer_counts(k, 0.1)
        qn = counts.into_normalized_l2()
        s, i = topx(qn)
        if sync_mode == "device":
            torch.cuda.synchronize(dev)
        return counts, s, i

    def step_e2e():
        m = api.MyrSeqs.from_packed(ctx, p_words, p_woff, p_boff, p_qual)      # H2D from pinned memory
        counts = m.compute_kmer_counts(k, 0.1)
        qn = counts.into_normalized_l2()
        s, i = topx(qn)
        if q_hi > q_lo:                                                        # D2H of this rank's share
            out_s_t[:q_hi - q_lo].copy_(s[q_lo:q_hi], non_blocking=True)
            out_i_t[:q_hi - q_lo].copy_(i[q_lo:q_hi], non_blocking=True)
        torch.cuda.synchronize(dev)
        return counts

    # ---- warm-up, then the timed region of the resident path; the e2e path is warmed up right
    # before its own timed region (switching paths reshuffles the memory pool once)
    for _ in range(args.warmup):
        step_resident()

    if rank == 0 and not os.environ.get("MYRIO_BENCH_NO_SMI"):  # diagnosis only: the line then carries no clocks
        sampler = ClockSampler(local_rank)
        if os.environ.get("MYRIO_BENCH_NVML_THREAD"):
            sampler.start()
        else:
            sampler.start_inline()
            sampler.poll()
    step_resident()  # one more warm-up step after the clock query (every rank: the step holds collectives)
    ctx.prof_reset()
    ctx.prof_enable(True)
    launches0 = ctx.launch_count
    ms_value = timed(step_resident, args.steps, "value")
    launches = ctx.launch_count - launches0
    k4_ms, k4_n = ctx.prof_query("k4_score")
    k1_ms, k1_n = ctx.prof_query("k1_")
    k5_ms, k5_n = ctx.prof_query("k5_")
    k4a_ms, _ = ctx.prof_query("k4a_")
    k7_ms, _ = ctx.prof_query("k7_")
    all_ms, all_n = ctx.prof_query("")
    ctx.prof_enable(False)
    ctx.prof_reset()
    for _ in range(max(1, min(args.warmup, 3))):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps, "e2e")
    clocks = sampler.stop() if sampler is not None else None
    sampler = None

    # ---- the same step with EVERY score finished (the fused top-x normally finishes only the scores whose
    # upper bound reaches the query's running bound; the top-x is bit-identical either way)
    ctx.set_topx_bound_pruning(False)
    step_resident()
    ms_all = timed(step_resident, max(1, min(args.steps, 2)))
    ctx.set_topx_bound_pruning(True)
    step_resident()

    # ---- k-mer counting THROUGH THE CALL (reads): myrio_count_reads + sync, wall clock, max over ranks
    n_cr = max(5, args.steps)
    c_tmp = resident.compute_kmer_counts(k, 0.1)  # warm-up
    read_nnz = c_tmp.nnz
    call_s = []
    for _ in range(n_cr):
        del c_tmp  # a caller holds one result at a time (as a step does)
        barrier()
        t0 = time.perf_counter()
        c_tmp = resident.compute_kmer_counts(k, 0.1)
        ctx.sync()
        call_s.append(time.perf_counter() - t0)
    del c_tmp
    t_count_reads_call = allmax(sorted(call_s)[len(call_s) // 2])  # median call, slowest rank

    # ---- exact algorithmic work of one step on this rank's shard (one untimed pass with the counters on)
    counts = resident.compute_kmer_counts(k, 0.1)
    qn = counts.into_normalized_l2()
    ts = np.zeros((Q, x), np.float32)
    ti = np.zeros((Q, x), np.uint32)
    _lib.check(L.myrio_score_topx(ctx.h, ttc.h, qn.h, 0, Q, x, api.SIM_COSINE, ts.ctypes.data_as(C.c_void_p),
                                  ti.ctypes.data_as(C.c_void_p)))
    st = ctx.score_stats()
    _, _, _, hck = counts.download()
    nq_sum = st["query_keys"]
    t_q = st["postings_touched"]
    read_kmers = float(hck.sum())

    n_gpu_scores = float(Q) * float(args.leaves) * world
    value = n_gpu_scores * args.steps / (ms_value * 1e-3)
    e2e_value = n_gpu_scores * args.steps / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel (K4).  The kernel is bound by warp-instruction ISSUE, not by HBM:
    # 98 % of the postings a query touches are bits of 128-byte dense-key bitmaps that stay in L2, so DRAM
    # moves a few % of its peak (ncu: `traffic`), while the SM sub-partitions issue near their limit.
    peak, peak_src = peaks()
    alg_bytes_step = 12 * nq_sum + 16 * nq_sum + 8 * t_q + 8 * x * Q
    k4_per_step = max(k4_n // max(args.steps, 1), 1)
    k4_avg_ms = k4_ms / max(k4_n, 1)
    k4_step_ms = k4_ms / max(args.steps, 1)
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    # the committed ncu capture of the CURRENT K4 (instructions per posting, DRAM bytes), older ones as fall-backs
    prof = latest_profile("r02_k4v12") or latest_profile("r02_k4v11") or latest_profile("r02_k4_ncu_summary")
    issue_peak = 148 * 4 * sm_mhz * 1e6  # warp instructions / s: 148 SMs x 4 schedulers x 1 issue / clk
    roofline = {
        "bound": "issue",
        "kernel": "k4_score_tiles_topx (score_tiles_count_kernel<COSINE, fused top-x>), seed pass + tile-major sweep",
        "unit": "Gwarp-inst/s", "peak": issue_peak / 1e9,
        "peak_source": f"148 SMs x 4 warp schedulers x SM clock under load ({sm_mhz:.0f} MHz, nvidia-smi during the timed region)",
        "launch_ms_avg": k4_avg_ms, "launches_per_step": k4_per_step, "kernel_ms_per_step": k4_step_ms,
        "kernel_share_of_step": k4_ms / max(all_ms, 1e-9),
        "postings_touched_per_step": t_q, "query_keys_per_step": nq_sum,
        "algorithmic_bytes_per_step_survey_8d": alg_bytes_step,
    }
    if prof:
        pname, pj = prof
        pl = pj.get("launches", [])
        wi = float(sum(l.get("warp_instructions", 0.0) for l in pl))
        dram = float(sum(l.get("dram_bytes", 0.0) for l in pl))
        # the round-1 capture predates the field: its workload touched 2.4347e11 postings per step (BENCH_r01)
        post_prof = float(pj.get("postings_touched_per_step") or (2.4347e11 if pname.startswith("r01_k4_v14") else 0.0))
        # warp instructions of THIS run = instructions per posting of the committed capture (same workload; the
        # capture's `state` names the commit it was taken at) x the postings counted in this run; / measured
        # kernel time / issue peak
        if wi > 0:
            wi_run = wi * (t_q / post_prof) if post_prof > 0 else wi
            roofline.update({
                "achieved": wi_run / (k4_step_ms * 1e-3) / 1e9,
                "frac": wi_run / (k4_step_ms * 1e-3) / issue_peak,
                "warp_instructions_per_step": wi_run,
                "thread_instructions_per_posting": 32.0 * wi / post_prof if post_prof > 0 else None,
                "instruction_count_source": f"profiles/{pname} (ncu smsp__inst_executed.sum of the K4 launches of one step"
                                            + (", scaled by postings touched" if post_prof > 0 else "") + ")",
            })
        if dram > 0:
            roofline["traffic"] = dram / max(len(pl), 1)
            roofline["traffic_source"] = f"profiles/{pname}: dram__bytes_read.sum + dram__bytes_write.sum per launch"
            roofline["hbm"] = {"achieved": dram / (k4_step_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                               "frac": dram / (k4_step_ms * 1e-3) / 1e9 / peak, "peak_source": peak_src,
                               "note": "measured DRAM bytes of the K4 launches of one step / measured kernel time"}
            roofline["l2_reuse_ratio"] = alg_bytes_step / dram
            roofline["l2_reuse_note"] = ("SURVEY §8d algorithmic bytes (8 B per posting touched) / DRAM bytes moved: posting "
                                         "lists and bitmaps are served from L2 and as bits, not as 8-byte postings")
    roofline.setdefault("achieved", None)
    roofline.setdefault("frac", None)
    roofline.setdefault("traffic", None)

    k1_bytes = k1_algorithmic_bytes(reads, read_nnz)
    k2_bytes = k2_algorithmic_bytes(tree, leaf_nnz)
    k1_step_ms = k1_ms / max(args.steps, 1)
    kmers = {
        "unit": "k-mers/s", "n_gpus": world,
        "reads": {
            "kmers_per_step": read_kmers,
            "k1_kernel": read_kmers / (k1_step_ms * 1e-3) if k1_ms > 0 else None,
            "count_call": read_kmers / t_count_reads_call,
            "count_call_ms": t_count_reads_call * 1e3, "k1_kernel_ms": k1_step_ms,
            "whole_job": read_kmers * world / t_count_reads_call,
            "note": "reads are replicated: every rank counts the whole batch (whole_job = N x count_call)",
            "roofline": {"bound": "hbm", "algorithmic_bytes": k1_bytes, "achieved": k1_bytes / (k1_step_ms * 1e-3) / 1e9 if k1_ms > 0 else None,
                         "peak": peak, "unit": "GB/s", "frac": k1_bytes / (k1_step_ms * 1e-3) / 1e9 / peak if k1_ms > 0 else None,
                         "bytes_rule": "SURVEY §8d K1: ceil(n/4) + n in, 12 D + 16 out per read"},
        },
        "leaves": {
            "kmers_logical": leaf_kmers,
            "k2_kernel_logical": leaf_kmers / (k2_ms * 1e-3) if k2_ms > 0 else None,
            "k2_kernel_reference_equivalent_x32": leaf_kmers * 32 / (k2_ms * 1e-3) if k2_ms > 0 else None,
            "count_call_logical": leaf_kmers / t_count_leaves,
            "count_call_ms": t_count_leaves * 1e3, "k2_kernel_ms": k2_ms,
            "whole_job_logical": allsum(leaf_kmers) / allmax(t_count_leaves),
            "note": "leaves are sharded: whole_job = all ranks' leaf k-mers / slowest rank's call",
            "roofline": {"bound": "hbm", "algorithmic_bytes": k2_bytes, "achieved": k2_bytes / (k2_ms * 1e-3) / 1e9 if k2_ms > 0 else None,
                         "peak": peak, "unit": "GB/s", "frac": k2_bytes / (k2_ms * 1e-3) / 1e9 / peak if k2_ms > 0 else None,
                         "bytes_rule": "SURVEY §8d K2: ceil(n/2) in, 10 D + 12 out per leaf"},
        },
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_value / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_rank * world,
                "d2h_bytes_per_step": d2h_total, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step_per_rank": h2d_rank,
                "note": "every rank uploads the read batch (queries are replicated) and downloads its 1/N of the merged top-x"},
        "gpu_launches": int(launches),
        "all_scores_finished": {
            "ms_per_step": ms_all / max(1, min(args.steps, 2)),
            "value": n_gpu_scores * max(1, min(args.steps, 2)) / (ms_all * 1e-3), "unit": UNIT,
            "note": "the same step with myrio_ctx_set_topx_bound_pruning(0): every (query, leaf) score is replayed to "
                    "its final f32 value instead of only those whose upper bound reaches the query's running x-th best "
                    "score; `value` counts every pair either way (each pair's shared keys are counted and its score "
                    "bounded), the top-x is bit-identical"},
        "roofline": roofline,
        "clocks": clocks,
        "kmers": kmers,
        "comm": comm_info,
        "step_ms_rank0": step_trace,
        "setup_s": {"leaves_fasta_ingest": t_upload_leaves, "k2_count_leaves": t_count_leaves, "k3_index_build": t_index},
        "index": info,
        "kernel_ms_per_step": {"k1_count_reads": k1_ms / args.steps, "k4_score_tiles_topx": k4_ms / args.steps,
                               "k4a_plan": k4a_ms / args.steps, "k5_topx_merge_tiles": k5_ms / args.steps,
                               "k7_merge": k7_ms / args.steps, "all_kernels": all_ms / args.steps},
    }

    # ---- N > 1: the in-library exchange and the plain torch.distributed exchange must agree (untimed)
    if world > 1:
        s1, i1 = mdist.sharded_topx(ctx, ttc, qn, x, api.SIM_COSINE, max_batch=args.topx_batch)
        s2, i2 = mdist.sharded_topx_allgather(ctx, ttc, qn, x, api.SIM_COSINE)
        torch.cuda.synchronize(dev)
        bad = int((~((s1.view(torch.int32) == s2.view(torch.int32)).all(dim=1) & (i1 == i2).all(dim=1))).sum().item())
        line["parity_in_bench"] = {"queries_checked": int(Q), "two_phase_vs_allgather_topx_mismatches": bad}

    # ---- stage 3 (clustering): not part of `value`, reported beside it; N > 1 shards the rows
    if not args.no_clustering:
        cpu_reads = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_reads = reads.subset(np.arange(min(reads.n, 4096)))
        line["clustering"] = bench_clustering(
            ctx, api, resident, args.clusters, (clocks or {}).get("sm_mhz") if clocks else None, cpu_reads,
            min(args.cpu_seconds, 5.0), world, (lambda m, prm: mdist.cluster_sharded(m, prm)) if world > 1 else None,
            allmax, allsum, args.cluster_reps, args.cluster_warmup)

    # ---- CPU baseline: the oracle on the host cores, rank 0, N=1 only, bounded sample
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        oracle.set_num_threads(host_threads())
        threads = oracle.max_threads()
        nleaf = min(tree.n, 2048)
        sub_tree = tree.subset(np.arange(nleaf))
        t0 = time.perf_counter()
        oracle.count_leaves(synth.tree_to_iupac(sub_tree), sub_tree.base_off, k, 32, threads=threads)
        t_cl = time.perf_counter() - t0
        sub_reads = reads.subset(np.arange(min(reads.n, 8192)))
        t0 = time.perf_counter()
        orc = oracle.count_reads(sub_reads.codes, sub_reads.qual, sub_reads.base_off, k, 0.1, threads=1)
        t_cr = time.perf_counter() - t0
        # leaf vectors for the score sample come from the GPU counts (already parity-checked in tests/)
        lro, lkeys, lvals, _ = store.kmer_store_counts_vec[0].download(aux=False)
        oln = oracle.normalize(oracle.Csr(lro, lkeys, lvals), threads)
        qro, qkeys, qvals, _ = qn.download(aux=False)
        oq = oracle.Csr(qro, qkeys, qvals)
        t0 = time.perf_counter()
        kq, vq = oq.row(0)
        oracle.score_all(oln, kq, vq, threads=threads)
        per_q = max(time.perf_counter() - t0, 1e-6)
        n_s = int(max(4, min(Q, args.cpu_seconds / per_q)))
        t0 = time.perf_counter()
        mism = 0
        for i in range(n_s):
            kq, vq = oq.row(i)
            s = oracle.score_all(oln, kq, vq, threads=threads)
            es, ei = oracle.topx(s, x, id_base=rank * args.leaves)
            mism += int((ei != ti[i]).any() or (es.view(np.uint32) != ts[i].view(np.uint32)).any())
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {
            "value": n_s * tree.n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {n_s} of {Q} queries x all {tree.n} leaves (merge-join dots + top-{x}), "
                      f"{threads} OpenMP threads over leaves like rayon",
            "kmers": {"leaves_reference_equivalent_x32": logical_leaf_kmers(sub_tree, k) * 32 / t_cl,
                      "leaves_logical": logical_leaf_kmers(sub_tree, k) / t_cl,
                      "reads_serial_like_the_reference": float(orc.aux.sum()) / t_cr,
                      "sample": f"{nleaf} leaves x 32 resamples on {threads} threads; {sub_reads.n} reads on 1 thread "
                                f"(clustering.rs:78-90 and run.rs:301-309 count reads sequentially)"},
        }
        line["parity_in_bench"] = {"queries_checked": n_s, "topx_mismatches": mism}
    del ttc, store, qn, counts

    # ---- strong scaling (configs[4]): the 1 M-leaf tree sharded over the ranks x one fixed read batch
    if args.strong_leaves > 0:
        line["strong"] = bench_strong(args, ctx, api, mdist, synth, rank, world, timed, allmax, dev)

    # ---- the sparse path (N = 1): where the dense-key bitmaps thin out
    if world == 1 and not args.no_variants:
        line["variants"] = bench_variants(args, ctx, api, synth, timed, tree0)

    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        ctx.comm_destroy()
        dist.destroy_process_group()


def bench_strong(args, ctx, api, mdist, synth, rank, world, timed, allmax, dev):
    """BASELINE.json configs[4] as a strong-scaling point: total work fixed (strong_reads x strong_leaves
    scores per step), the tree ALWAYS the same 8 sub-trees, rank r holding sub-trees [8r/N, 8(r+1)/N); reads
    drawn from every sub-tree (12.5 k each), the same batch at every N."""
    if N_STRONG_SUBTREES % world != 0:
        return {"skipped": f"{N_STRONG_SUBTREES} sub-trees do not split over {world} ranks"}
    per_rank = N_STRONG_SUBTREES // world
    per_sub = args.strong_leaves // N_STRONG_SUBTREES
    t0 = time.perf_counter()
    subs, read_parts = [], []
    for j in range(rank * per_rank, (rank + 1) * per_rank):
        sub = strong_subtree(args, j)
        subs.append(sub)
        read_parts.append(strong_reads_of(args, sub, j))
    tree = synth.concat(subs)
    local_reads = synth.concat(read_parts)
    reads = mdist.allgather_reads(local_reads) if world > 1 else local_reads  # rank order == sub-tree order
    t_gen = time.perf_counter() - t0
    k, x = args.k, args.x
    text, text_off = ascii_leaves(tree)
    t0 = time.perf_counter()
    store, _ = api.TaxTreeStore.from_fasta_records(ctx, text, text_off, leaf_id_base=rank * per_rank * per_sub)
    store.compute_and_append_kmer_counts(k, 32)
    ttc = api.TaxTreeCompute.from_store_tree(store, k, 32, tile_leaves=args.tile_leaves)
    store.kmer_store_counts_vec, store.k_precomputed = [], []
    ctx.sync()
    t_build = allmax(time.perf_counter() - t0)
    resident = api.MyrSeqs(ctx, reads)
    Q = reads.n

    import torch
    res = (torch.empty((Q, x), dtype=torch.float32, device=dev), torch.empty((Q, x), dtype=torch.int32, device=dev))
    torch.cuda.synchronize(dev)

    def step():
        qn = resident.compute_kmer_counts(k, 0.1).into_normalized_l2()
        if world > 1:
            return mdist.sharded_topx(ctx, ttc, qn, x, api.SIM_COSINE, max_batch=args.topx_batch, out=res)
        return mdist.sharded_topx_allgather(ctx, ttc, qn, x, api.SIM_COSINE, out=res)

    step()
    ms = timed(step, args.strong_steps)
    scores = float(Q) * float(per_sub * N_STRONG_SUBTREES)
    out = {"workload": f"configs[4]: 1 tree of {per_sub * N_STRONG_SUBTREES} leaves ({N_STRONG_SUBTREES} fixed sub-trees, seed 1005) "
                       f"sharded over {world} GPU(s) x {Q} reads (seed 2005+j per sub-tree), k={k}, top-{x}",
           "scaling": "strong", "leaves_total": per_sub * N_STRONG_SUBTREES, "reads": Q, "n_gpus": world,
           "steps": args.strong_steps, "ms_per_step": ms / args.strong_steps,
           "value": scores * args.strong_steps / (ms * 1e-3), "unit": UNIT,
           "index": ttc.info(), "setup_s": {"generate": t_gen, "ingest_count_index_max_over_ranks": t_build}}
    del ttc, store, resident
    return out


def bench_variants(args, ctx, api, synth, timed, tree0):
    """Scoring throughput where the conserved-block bitmaps do not carry the load (extra keys, not the headline):
    (a) the same tree generator WITHOUT the 160-bp conserved block; (b) the headline tree with 1 % of the bases
    replaced by ambiguity codes (resampled counts are no longer the common count -> dense keys demoted to sparse
    postings); (c) the headline tree at k = 12 (short k-mers: long sparse posting lists)."""
    out = {}
    nq = min(args.variant_reads, args.reads)

    def run(tag, store, reads, k, desc):
        ttc = api.TaxTreeCompute.from_store_tree(store, k, 32, tile_leaves=args.tile_leaves)
        store.kmer_store_counts_vec, store.k_precomputed = [], []
        resident = api.MyrSeqs(ctx, reads)
        qn = resident.compute_kmer_counts(k, 0.1).into_normalized_l2()
        ts = np.zeros((reads.n, args.x), np.float32)
        ti = np.zeros((reads.n, args.x), np.uint32)
        from myrio_b200 import _lib
        _lib.check(ctx._L.myrio_score_topx(ctx.h, ttc.h, qn.h, 0, reads.n, args.x, api.SIM_COSINE,
                                           ts.ctypes.data_as(C.c_void_p), ti.ctypes.data_as(C.c_void_p)))
        st = ctx.score_stats()

        def step():
            q2 = resident.compute_kmer_counts(k, 0.1).into_normalized_l2()
            api.TaxTreeResults.from_compute_tree(q2, ttc, api.SIM_COSINE, args.x)

        step()
        ctx.prof_reset()
        ctx.prof_enable(True)
        ms = timed(step, 2)
        k4_ms, _ = ctx.prof_query("k4_score")
        ctx.prof_enable(False)
        ctx.prof_reset()
        info = ttc.info()
        out[tag] = {"workload": desc, "leaves": info["n_leaves"], "reads": reads.n, "k": k,
                    "ms_per_step": ms / 2, "value": reads.n * info["n_leaves"] * 2 / (ms * 1e-3), "unit": UNIT,
                    "k4_kernel_ms_per_step": k4_ms / 2, "postings_touched_per_step": st["postings_touched"],
                    "postings_per_score": st["postings_touched"] / max(1.0, float(reads.n) * info["n_leaves"]),
                    "index": info, "note": "value includes the D2H of the top-x (public call); 2 timed steps"}

    # (a) no conserved block
    t_a = synth.make_tree(args.leaves, 1012, root_seed=1012, conserved_len=0)
    r_a = synth.make_reads(t_a, nq, 2012)
    text, off = ascii_leaves(t_a)
    st_a, _ = api.TaxTreeStore.from_fasta_records(ctx, text, off)
    run("no_conserved_block", st_a, r_a, args.k, f"tree seed 1012 without the 160-bp conserved block, {nq} of its reads")
    del st_a, t_a
    # (b) 1 % ambiguity codes in the headline tree
    rng = np.random.default_rng(77)
    iupac = synth.tree_to_iupac(tree0).copy()
    pos = rng.choice(len(iupac), len(iupac) // 100, replace=False)
    iupac[pos] = rng.choice(np.array([3, 5, 6, 9, 10, 12, 7, 11, 13, 14, 15], np.uint8), len(pos))
    st_b, _ = api.TaxTreeStore.from_fasta_records(ctx, synth.iupac_to_ascii(iupac), np.ascontiguousarray(tree0.base_off, np.uint64))
    r_b = synth.make_reads(tree0, nq, 2002)
    run("ambiguity_1pct", st_b, r_b, args.k, f"headline tree with 1 % of the bases replaced by IUPAC ambiguity codes "
                                             f"(32 resamples on the device), {nq} reads")
    del st_b
    # (c) k = 12
    text, off = ascii_leaves(tree0)
    st_c, _ = api.TaxTreeStore.from_fasta_records(ctx, text, off)
    run("k12", st_c, r_b, 12, f"headline tree at k = 12, {nq} reads")
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
