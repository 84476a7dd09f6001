#!/usr/bin/env python
"""Headline benchmark: cosine top-100 retrieval, 10k queries over a 1M x 2048-d
database (BASELINE.json configs[3], the configuration `metric` is quoted on; it
fits one B200: 8 GB fp32 + 4 GB bf16).

    python bench.py --gpus N --steps K --warmup W            # this framework
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path

A step = one pass of the hot path over one batch of synthetic queries: bf16
tcgen05 screen with fused streaming top-k, exact fp64-accumulated re-rank, and
for N > 1 (database row-sharded, one process per GPU) the candidate exchange:
an NCCL all-gather of the shards' candidate screen scores, the global
threshold, an exact re-rank of each shard's own candidates above it, an
all-gather of the per-shard lists and the certified merge.  Prints ONE JSON line on rank 0.
"""

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "queries/s top-100 over 1Mx2048-d DB"
UNIT = "queries/s"
Q_DEFAULT, N_DEFAULT, D_DEFAULT, K_DEFAULT = 10000, 1000000, 2048, 100
SEED = 1234 + 4  # manual_seed(1234 + cfg)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--queries", type=int, default=Q_DEFAULT)
    ap.add_argument("--db-rows", type=int, default=N_DEFAULT)
    ap.add_argument("--dim", type=int, default=D_DEFAULT)
    ap.add_argument("--k", type=int, default=K_DEFAULT)
    ap.add_argument("--cpu-sample-queries", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-upload", default="auto", choices=["auto", "split", "full"],
                    help="N > 1: how the pinned host queries reach every rank -- split: 1/N of the rows per rank + an "
                         "NVLink all-gather (on the search stream); full: every rank copies all rows over its own "
                         "PCIe link (copy engine only).  Measured on 8 x B200: split 1.91 M q/s, full 1.67 M q/s (eight "
                         "82 MB copies per step saturate the host side); auto = split")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the side measurements: region descriptors / mining (N=1 only) and the "
                         "end-to-end configs[4] leg (every N)")
    ap.add_argument("--e2e-db-rows", type=int, default=10000000)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    # B200_PROFILING.md fallback
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------ synthetic data
def make_rows(n, d, seed, device, chunk=131072):
    """normalize(randn(n, d)) rows, generated in chunks with a seeded generator.
    SURVEY.md 8d: continuous Gaussians => no exact ties."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((n, d), dtype=torch.float32, device=device)
    for s in range(0, n, chunk):
        x = torch.randn((min(chunk, n - s), d), generator=g, device=device)
        out[s:s + chunk] = x / x.norm(dim=1, keepdim=True)
    return out


class ClockSampler(object):
    """SM clock / power / throttle reasons DURING the timed region: NVML polled every
    2 ms from a thread (a timed region of a few tens of ms still gets samples);
    nvidia-smi -lms 100 when NVML cannot be opened."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    # nvmlClocksEventReason* bits
    REASON_BITS = (("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20),
                   ("hw_thermal_slowdown", 0x40), ("hw_power_brake_slowdown", 0x80))

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.nvml, self.handle, self.stop_flag, self.t = None, None, threading.Event(), None
        self.samples, self.mask, self.sm_max = [], 0, None

    def _open_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        handle = None
        try:
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            handle = None
        if handle is None:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.index
            if vis:
                ent = vis.split(",")[self.index].strip()
                if ent.isdigit():
                    phys = int(ent)
            handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
        self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
        pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)   # fails here rather than in the thread
        self.nvml, self.handle = pynvml, handle

    def _poll_nvml(self):
        nv, h = self.nvml, self.handle
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self.stop_flag.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.samples.append((sm, pw))
                if reasons is not None:
                    self.mask |= int(reasons(h))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            self._open_nvml()
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    @staticmethod
    def _summary(sm, pw, sm_max, reasons, source):
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": sm_max, "reasons": ["no samples"], "source": source}
        busy = [c for c, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": sm_max, "samples": len(sm),
                "power_w_max": max(pw), "reasons": sorted(reasons), "source": source}

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.t.join(timeout=2)
            reasons = [name for name, bit in self.REASON_BITS if self.mask & bit]
            return self._summary([s for s, _ in self.samples], [p for _, p in self.samples], self.sm_max,
                                 reasons, "nvml, 2 ms poll")
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])), mx.append(float(r[2])), pw.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return self._summary(sm, pw, max(mx) if mx else None, reasons, "nvidia-smi -lms 100")


# ------------------------------------------------------------------ reference arm
def cpu_topk(q, db, k, chunk=128):
    """The reference's CPU path for this metric, via the oracle's restatement:
    sim = torch.mm(q_chunk, db.t()) (test/siamese_regions_test.py:76) then the
    best k of every row (utils/metrics.py:11,33).  torch fp32 on all host cores."""
    import oracle  # the one place bench.py may execute oracle/ (cpu baseline)
    out_s, out_i = [], []
    for s in range(0, q.size(0), chunk):
        sim = oracle.similarity(q[s:s + chunk], db)
        v, i = sim.topk(k, dim=1)  # == first k of the descending sort, without the full sort
        out_s.append(v), out_i.append(i)
    return torch.cat(out_s), torch.cat(out_i)


def time_cpu(q, db, k, repeats=1):
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        cpu_topk(q, db, k)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def topk_parity(q, db, k, gpu_scores, gpu_idx, cpu_scores, cpu_idx, tie_rtol=16 * 2.0 ** -24, tie_atol=5e-8):
    """GPU top-k of a query sample against the CPU oracle's (torch fp32 mm + topk on the host).
    rows identical to the oracle are index-exact; a row that differs is adjudicated in fp64 on
    the union of the two lists: it counts as fp64_adjudicated when the fp64 ranking of that
    union is the GPU's list AND every differing entry is an fp32-noise tie in the oracle's own
    scores (gap <= tie_atol + tie_rtol |score|: 16 ulp); anything else is unexplained."""
    gs, gi = gpu_scores.cpu(), gpu_idx.cpu()
    rows = gi.size(0)
    same = (gi == cpu_idx).all(dim=1)
    adjudicated, unexplained, max_gap, entries = 0, 0, 0.0, 0
    for r in (~same).nonzero().flatten().tolist():
        entries += int((gi[r] != cpu_idx[r]).sum())
        union = torch.unique(torch.cat([gi[r], cpu_idx[r]]))          # sorted: ties -> lower index
        s64 = db[union].double() @ q[r].double()
        order = torch.sort(s64, descending=True, stable=True).indices[:k]
        pos = (gi[r] != cpu_idx[r]).nonzero().flatten()
        gaps = ((db[gi[r, pos]] * q[r]).sum(1) - cpu_scores[r, pos]).abs()
        max_gap = max(max_gap, gaps.max().item())
        if torch.equal(union[order], gi[r]) and bool((gaps <= tie_atol + tie_rtol * cpu_scores[r, pos].abs()).all()):
            adjudicated += 1
        else:
            unexplained += 1
    rel = ((gs - cpu_scores).abs() / cpu_scores.abs().clamp_min(1e-6)).max().item()
    return {"rows": rows, "index_exact_rows": int(same.sum()), "fp64_adjudicated": adjudicated,
            "unexplained_rows": unexplained, "entries_differing": entries, "max_tie_gap": max_gap,
            "tie_tolerance": "%.1e + %.1e |score|" % (tie_atol, tie_rtol), "max_rel_score_err": rel, "k": k,
            "oracle": "torch fp32 mm + topk on the host (oracle.similarity), first %d queries over the "
                      "full database" % rows}


def run_reference(a, rank, world):
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    torch.set_num_threads(os.cpu_count() or 1)
    cores = torch.get_num_threads()
    g = torch.Generator().manual_seed(SEED)
    # same distribution as the GPU arm; generated on the host in chunks
    db = torch.empty((a.db_rows, a.dim), dtype=torch.float32)
    for s in range(0, a.db_rows, 65536):
        x = torch.randn((min(65536, a.db_rows - s), a.dim), generator=g)
        db[s:s + 65536] = x / x.norm(dim=1, keepdim=True)
    nq = a.cpu_sample_queries
    qs = torch.randn((nq, a.dim), generator=g)
    qs = qs / qs.norm(dim=1, keepdim=True)
    for _ in range(max(1, min(a.warmup, 1))):
        cpu_topk(qs[:32], db, a.k)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        cpu_topk(qs, db, a.k)
    dt = (time.perf_counter() - t0) / a.steps
    qps = nq / dt
    sample = "%d of %d queries per step over the full %d-row database, torch fp32 mm + topk(%d), chunks of 128" % (
        nq, a.queries, a.db_rows, a.k)
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, max(1, world)),
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(a, world):
    return {"workload": "cosine top-%d retrieval, %d queries x %d-row x %d-d database (BASELINE configs[3])" %
                        (a.k, a.queries, a.db_rows, a.dim),
            "queries": a.queries, "db_rows": a.db_rows, "dim": a.dim, "k": a.k,
            "db_sharding": "row-wise over %d GPU(s)" % world,
            "l2_policy": "inputs larger than L2 (bf16 database shard %.2f GB > 126 MB)" %
                         (a.db_rows / world * a.dim * 2 / 1e9)}


# ------------------------------------------------------------------ B200 arm
def run_b200(a, rank, world, local_rank):
    import torch.distributed as dist
    from instance_search_b200.search import ShardedIndex, shard_bounds
    from instance_search_b200 import _lib

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.lib().isb_check_device(), "isb_check_device")  # no fallback

    lo, hi = shard_bounds(a.db_rows, world)[rank]
    # every rank draws the same global stream and keeps its rows (shards of ONE database)
    shard = make_rows_slice(a.db_rows, a.dim, SEED, dev, lo, hi)
    index = ShardedIndex(shard, a.db_rows, rank, world)
    del shard
    q_dev = make_rows(a.queries, a.dim, SEED + 100, dev)
    q_host = q_dev.cpu().pin_memory()
    out_s_host = torch.empty((a.queries, a.k), dtype=torch.float32).pin_memory()
    out_i_host = torch.empty((a.queries, a.k), dtype=torch.int64).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()          # tickets resolved / last results landed: still inside the timed region
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    screen_events = []
    pending = []          # exactness tickets of queued steps (ops.ExactnessTicket)

    def step_resident():
        # the certificate's 4-byte read-back is deferred: step i's is checked after step i + 1
        # has been queued, so no step waits on a host round trip (every ticket is resolved
        # inside the timed region, see drain)
        s, i, t = index.search(q_dev, a.k, events=screen_events, defer=True)
        pending.append(t)
        if len(pending) > 1:
            t0 = pending.pop(0)
            if t0 is not None:
                t0.resolve()

    def drain():
        while pending:
            t0 = pending.pop(0)
            if t0 is not None:
                t0.resolve()

    # ---- end to end from host buffers, double-buffered like a serving loop: the H2D copy of batch
    # i + 1 (copy stream) and the D2H copy of batch i - 1's results (copy-out stream) overlap the
    # search of batch i; every batch's copies, and the wait for its results, are inside the timed region
    copy_in, copy_out = torch.cuda.Stream(), torch.cuda.Stream()
    inflight = {"q": None, "ready": None, "out": []}

    upload_full = world > 1 and a.e2e_upload == "full"

    def start_upload():
        with torch.cuda.stream(copy_in):
            lo_q, hi_q = shard_bounds(a.queries, world)[rank] if (world > 1 and not upload_full) else (0, a.queries)
            part = q_host[lo_q:hi_q].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        inflight["q"], inflight["ready"] = part, ev

    def step_e2e():
        if inflight["q"] is None:
            start_upload()                         # first batch of the timed region
        torch.cuda.current_stream().wait_event(inflight["ready"])
        part = inflight["q"]
        part.record_stream(torch.cuda.current_stream())
        q = index.gather_queries(part, a.queries) if (world > 1 and not upload_full) else part
        start_upload()                             # next batch's H2D overlaps this batch's search
        s, i, t = index.search(q, a.k, defer=True)
        done = torch.cuda.Event()
        done.record()
        with torch.cuda.stream(copy_out):
            copy_out.wait_event(done)
            out_s_host.copy_(s, non_blocking=True)
            out_i_host.copy_(i, non_blocking=True)
            s.record_stream(copy_out), i.record_stream(copy_out)
            landed = torch.cuda.Event()
            landed.record()
        inflight["out"].append((t, landed))
        if len(inflight["out"]) > 1:               # the caller takes batch i - 1's results now
            t0, l0 = inflight["out"].pop(0)
            if t0 is not None:
                t0.resolve()
            l0.synchronize()

    def drain_e2e():
        while inflight["out"]:
            t0, l0 = inflight["out"].pop(0)
            if t0 is not None:
                t0.resolve()
            l0.synchronize()
        inflight["q"] = None

    for _ in range(a.warmup):
        step_resident()
    drain()
    screen_events.clear()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_resident, a.steps, drain)
    clocks = sampler.stop() if rank == 0 else None
    ev = list(screen_events)
    screen_ms = sum(e0.elapsed_time(e1) for e0, e1 in ev) / max(1, len(ev))
    for _ in range(min(2, a.warmup)):
        step_e2e()
    drain_e2e()
    ms_e2e = timed(step_e2e, a.steps, drain_e2e)

    ms_step = ms_total / a.steps
    value = a.queries / (ms_step * 1e-3)
    e2e_value = a.queries / (ms_e2e / a.steps * 1e-3)

    # ---- roofline of the dominant kernel: gemm_tc_kernel<TopkSched, TopkEpilogue>
    pk, pk_src = peaks()
    rows_local = hi - lo
    flops = 2.0 * a.queries * rows_local * a.dim          # 2*Q*N*D per launch
    achieved = flops / (screen_ms * 1e-3) / 1e12
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])  # timed inside a long step
    traffic, traffic_src = None, None
    if a.queries == Q_DEFAULT and a.dim == D_DEFAULT and a.k == K_DEFAULT:
        traffic, traffic_src = profile_traffic(rows_local)
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes": 2.0 * rows_local * a.dim + 2.0 * a.queries * a.dim + 8.0 * a.queries * a.k,
                "kernel": "gemm_tc_pair_kernel<TopkSched,TopkEpilogue> (tcgen05 cta_group::2 screen, 256x256 tiles, "
                          "streaming top-k epilogue)",
                "kernel_ms": screen_ms, "peak_source": pk_src + " bf16_tflops_sustained",
                "frac_of_burst_peak": achieved / pk["bf16_tflops"]}

    # the result every rank holds after the timed steps (identical on all ranks), for the parity check
    res_s, res_i = index.search(q_dev, a.k)
    torch.cuda.synchronize()
    # the same kernel with nothing running beside it (in the timed region the exact re-rank / candidate
    # exchange of the previous batch runs on a side stream under it and takes part of the SMs and of the
    # power budget): three plain searches, outside the timed region, for the record only
    alone_events = []
    for _ in range(3):
        index.search(q_dev, a.k, events=alone_events)
    torch.cuda.synchronize()
    alone_ms = min(e0.elapsed_time(e1) for e0, e1 in alone_events) if alone_events else None
    if alone_ms:
        roofline["kernel_ms_alone"] = alone_ms
        roofline["frac_alone"] = flops / (alone_ms * 1e-3) / 1e12 / peak
        roofline["note"] = ("achieved / frac: the screen as timed inside the steps, with the previous batch's tail "
                            "running concurrently on a side stream; *_alone: the same launch with nothing beside it")

    line = None
    if rank == 0:
        cpu, parity = None, None
        if not a.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            nq = min(a.cpu_sample_queries, a.queries)
            if world == 1:
                db_host = index.local.db_f32[:, :a.dim].cpu()
            else:   # rank 0 draws the whole database (the same global stream) and keeps it on the host
                db_host = make_rows_host(a.db_rows, a.dim, SEED, dev)
            t0 = time.perf_counter()
            cpu_s, cpu_i = cpu_topk(q_host[:nq], db_host, a.k)
            dt = time.perf_counter() - t0
            if world == 1:
                cpu = {"value": nq / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                       "sample": "first %d of %d queries over the full %d-row database (%.1f s), torch fp32 "
                                 "mm + topk(%d) on the host" % (nq, a.queries, a.db_rows, dt, a.k)}
            parity = topk_parity(q_host[:nq], db_host, a.k, res_s[:nq], res_i[:nq], cpu_s, cpu_i)
            del db_host
        # N = 1: bf16 cast, thr init, screen, rerank.  N > 1: bf16 cast, thr init, screen, candidates,
        # global threshold, rerank of the owned candidates, certified merge (NCCL's kernels not counted)
        launches_per_step = 4 if world == 1 else 7
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16 screen + f64-accumulated f32 re-rank",
            "data": "synthetic (seeded unit-norm Gaussian rows)", "config": workload_config(a, world),
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity,
            "e2e": {"value": e2e_value, "unit": UNIT,
                    # query bytes that cross PCIe per step over the whole job in split mode (1/N of the rows per rank);
                    # in full mode every rank copies all of them
                    "h2d_bytes_per_step": q_host.numel() * 4,
                    "h2d_mode": ("full rows on every rank over its own PCIe link" if upload_full else
                                 "1/N of the rows per rank + NVLink all-gather") if world > 1 else "single GPU",
                    "d2h_bytes_per_step": out_s_host.numel() * 4 + out_i_host.numel() * 8,
                    "ms_per_step": ms_e2e / a.steps,
                    "pipeline": "double-buffered serving loop: pinned-host queries of batch i+1 are copied H2D "
                                "(see h2d_mode) and the results of batch "
                                "i-1 D2H on copy streams while batch i is searched; all copies of all %d batches "
                                "and the wait for the last results are inside the timed region" % a.steps},
            "gpu_launches": launches_per_step * a.steps * 2,  # resident + e2e timed regions
            "clocks": clocks,
            # rows searched / re-screened fp32-grade / searched exhaustively (N > 1: rows the global
            # certificate sent back to the shards' own certified search)
            "exactness": dict(index.local.stats) if world == 1 else dict(index.stats),
        }
        if world == 1 and not a.no_secondary:
            del index
            torch.cuda.empty_cache()
            line["secondary"] = {"region_descriptors": side_regions(dev, pk, not a.no_cpu_baseline),
                                 "mining": side_mining(dev, pk, not a.no_cpu_baseline)}
    if not a.no_secondary:
        index = None
        torch.cuda.empty_cache()
        e2e = side_e2e(a, dev, rank, world, not a.no_cpu_baseline)      # collective: every rank runs it
        if line is not None:
            line.setdefault("secondary", {})["e2e_instance_search"] = e2e
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ side measurements (N = 1)
# BASELINE.json's metric also names "region descriptors/s" (configs[1]) and the path includes
# the mining of configs[2]; they are measured here after the headline, on rank 0, a few ms each.
SCREEN_PROFILES = {   # rows of the screened shard -> committed ncu --set full summary of the screen kernel
    1000000: ("profiles/r02_ncu_search_screen_1M.txt", "profiles/r01_ncu_search_screen_pair.txt"),
    125000: ("profiles/r02_ncu_search_screen_125k_shard.txt", "profiles/r01_ncu_search_screen_pair_125k_shard.txt"),
}


def profile_traffic(rows_local):
    """dram__bytes_read.sum + dram__bytes_write.sum of the screen kernel, per launch, parsed from the
    committed ncu summary of the same shape (a profile, not a measurement of this run: ncu cannot run
    inside the timed bench).  (None, None) when no capture of this shard size is committed."""
    for rel in SCREEN_PROFILES.get(int(rows_local), ()):
        path = os.path.join(ROOT, rel)
        if not os.path.exists(path):
            continue
        total, seen = 0.0, set()
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        with open(path) as f:
            for ln in f:
                parts = ln.split()
                if len(parts) == 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") \
                        and parts[0] not in seen:      # first launch listed in the summary
                    seen.add(parts[0])
                    total += float(parts[1]) * scale.get(parts[2], 1.0)
        if len(seen) == 2:
            return total, "from profile %s (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)" % rel
    return None, None


def _median_ms(fn, iters=10, warmup=3, flush=None):
    ts = []
    for it in range(warmup + iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if it >= warmup:
            ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def cpu_regions(x_host, host_w, k):
    """The reference's CPU path for region descriptors: forward_single image by image
    (model/siamese.py:185-223 via train/siamese_regions.py:31-38), oracle restatement,
    torch fp32 on all host cores.  Returns (seconds, desc, idx)."""
    import oracle
    t0 = time.perf_counter()
    d, _, i, _ = oracle.region_descriptor_forward(x_host, *host_w, k, (7, 7))
    return time.perf_counter() - t0, d, i


def side_regions(dev, pk, cpu_baseline=True):
    """configs[1]: 256 x 2048 x {14x14, 32x32} fp32 maps -> descriptors (eval path, D=2048, k=6)."""
    from instance_search_b200 import regions
    g = torch.Generator(device=dev).manual_seed(1234 + 2)
    B, C, ncls, D, k = 256, 2048, 464, 2048, 6
    Kin = C * 49
    lin_w_f32 = torch.randn(D, Kin, device=dev, generator=g) / Kin ** 0.5
    hw = regions.HeadWeights(torch.randn(ncls, C, device=dev, generator=g) / C ** 0.5,
                             0.01 * torch.randn(ncls, device=dev, generator=g),
                             0.01 * torch.randn(Kin, device=dev, generator=g),
                             lin_w_f32,
                             0.01 * torch.randn(D, device=dev, generator=g))
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)   # 256 MB > L2
    host_w = None
    if cpu_baseline:
        host_w = [t.cpu() for t in (hw.cls_w, hw.cls_b, hw.shift, lin_w_f32, hw.lin_b)]
    del lin_w_f32
    out = {}
    for hwsize in (14, 32):
        x = torch.relu(torch.randn(B, C, hwsize, hwsize, device=dev, generator=g))
        stats = {}
        ms = _median_ms(lambda: regions.region_descriptors(x, hw, k, (7, 7), want_cls_out=False, stats=stats),
                        flush=flush)
        ms_head = _median_ms(lambda: regions.region_head(x, hw, k, (7, 7), want_cls_out=False), flush=flush)
        # the same chain (+ projection) replayed from ONE CUDA graph (regions.GraphedRegionDescriptors)
        graphed = regions.GraphedRegionDescriptors(x, hw, k, (7, 7))
        ms_graph = _median_ms(graphed.replay, flush=flush)
        del graphed
        # the pooling pass alone (north_star: >= 70 % of HBM bandwidth): x read once, window means
        # written as bf16 hi + lo; CUDA events around the one kernel, L2 flushed before every launch
        ms_pool = _median_ms(lambda: regions.region_pool_probe(x, hw, k + regions.RUNNER_UPS, (7, 7)), flush=flush)
        rd, wr = regions.region_pool_probe(x, hw, k + regions.RUNNER_UPS, (7, 7))
        # the gather pass alone: crops of the k + 2 scored windows read, U = sum crop/|crop| + n shift
        # written as bf16 hi + lo, window means as a by-product
        ke = k + regions.RUNNER_UPS
        idx_e, nsel_e, _, norm_e, _, _, _ = regions.region_select(x, hw, ke, (7, 7))
        ms_gather = _median_ms(lambda: regions.region_gather(x, hw, ke, (7, 7), idx_e, nsel_e, norm_e, k_sum=k),
                               flush=flush)
        win_rows = idx_e.clamp(min=0) // (hwsize - 6)
        rows_read = (win_rows.max(1).values + 7 - win_rows.min(1).values).clamp(max=hwsize).double().mean().item()
        # bytes the kernel touches: small maps stage whole planes, larger ones the row range of the windows
        g_rd = 4 * B * C * hwsize * (hwsize if hwsize * hwsize <= 256 else rows_read)
        g_wr = 2 * 2 * B * Kin + 4 * B * ke * C
        # streamed: 8 batches queued back to back over two alternating inputs (each larger than
        # L2, so no batch finds its map cached), certificates read once at the end -- what a
        # pipelined embedding loop sees once launch latency is hidden
        x2 = torch.relu(torch.randn(B, C, hwsize, hwsize, device=dev, generator=g))
        nstream, pending = 8, []

        def stream():
            pending.clear()
            for i in range(nstream):
                pending.append(regions.region_descriptors_async((x, x2)[i & 1], hw, k, (7, 7),
                                                                want_cls_out=False)[4][:, 0])
        ms_stream = _median_ms(stream, iters=5, warmup=2) / nstream
        n_unc_stream = int(torch.stack(pending).sum().item())
        del x2
        regions._PROBE_CACHE.clear()
        cpu, parity = None, None
        if host_w is not None:
            # the reference's per-image CPU path on a bounded sample of the same batch, and the GPU
            # descriptors of those images against it
            ns = 32 if hwsize == 14 else 16
            torch.set_num_threads(os.cpu_count() or 1)
            dt, od, oi = cpu_regions(x[:ns].cpu(), host_w, k)
            gd, _, gi, _ = regions.region_descriptors(x, hw, k, (7, 7), want_cls_out=False)
            gd, gi = gd[:ns].cpu().double(), gi[:ns].cpu()
            u_cpu = ns * min((hwsize - 6) ** 2, k)
            cpu = {"value": u_cpu / dt, "unit": "region descriptors/s", "cores": torch.get_num_threads(),
                   "kind": "port", "sample": "first %d of %d images, forward_single looped per image "
                   "(oracle, torch fp32), %.1f s" % (ns, B, dt)}
            l2 = (gd - od.double()).norm(dim=1)
            parity = {"images": ns, "windows_identical_images": int((gi == oi).all(dim=1).sum()),
                      "max_descriptor_l2_err": float(l2.max()),
                      "max_1_minus_cos": float((1 - (gd * od.double()).sum(1) / (gd.norm(dim=1) * od.double().norm(dim=1))).max())}
        # SURVEY 8d bytes of the bandwidth-bound part: x once + classifier + the bf16 hi+lo operand
        nbytes = 4 * B * C * hwsize * hwsize + 4 * ncls * C + 2 * 2 * B * Kin
        units = B * min((hwsize - 6) ** 2, k)
        out["%dx%d" % (hwsize, hwsize)] = {
            "region_descriptors_per_s": units / (ms * 1e-3), "images_per_s": B / (ms * 1e-3), "ms_per_batch": ms,
            "pool_select_gather": {"bound": "hbm", "ms": ms_head, "algorithmic_bytes": nbytes,
                                   "achieved": nbytes / (ms_head * 1e-3) / 1e9, "peak": pk["hbm_gbs"],
                                   "unit": "GB/s", "frac": nbytes / (ms_head * 1e-3) / 1e9 / pk["hbm_gbs"]},
            "pooling_kernel": {"kernel": "region_pool_fast_kernel", "bound": "hbm", "ms": ms_pool,
                               "algorithmic_bytes": rd + wr, "achieved": (rd + wr) / (ms_pool * 1e-3) / 1e9,
                               "peak": pk["hbm_gbs"], "unit": "GB/s",
                               "frac": (rd + wr) / (ms_pool * 1e-3) / 1e9 / pk["hbm_gbs"]},
            "gather_kernel": {"kernel": "region_gather7_kernel" if hwsize * hwsize <= 256 else "region_gather_kernel",
                              "bound": "hbm", "ms": ms_gather, "bytes_touched": g_rd + g_wr,
                              "achieved": (g_rd + g_wr) / (ms_gather * 1e-3) / 1e9, "peak": pk["hbm_gbs"],
                              "unit": "GB/s", "frac": (g_rd + g_wr) / (ms_gather * 1e-3) / 1e9 / pk["hbm_gbs"]},
            "cuda_graph_replay": {"ms_per_batch": ms_graph, "region_descriptors_per_s": units / (ms_graph * 1e-3),
                                  "what": "head + projection captured once (regions.GraphedRegionDescriptors), one "
                                          "graph launch per batch; certificates read by the caller"},
            "projection_ms": ms - ms_head,
            "streamed": {"ms_per_batch": ms_stream, "region_descriptors_per_s": units / (ms_stream * 1e-3),
                         "batches_in_flight": nstream, "uncertified_images_last_pass": n_unc_stream},
            "batches_resolved_exactly": stats.get("batches_resolved_exactly", 0),
            "images_resolved_exactly": stats.get("images_resolved_exactly", 0), "batches": stats.get("batches", 0),
            "cpu_baseline": cpu, "parity": parity}
        del x
    return {"workload": "region descriptors (eval), 256 x 2048 x HxW fp32 maps, ncls=464, k=6, D=2048 "
                        "(BASELINE configs[1]); L2 flushed between batches", **out}


def cpu_mining(E, lab, anchors, positives, semi, n_couples):
    """The reference's CPU path for mining: similarities = mm(E, E.t()) (utils/train_siamese.py:53),
    then per couple the masked arg-max of train/siamese_regions.py:106-126 (oracle restatement).  The
    product is timed in full; the per-couple selection on the first n_couples couples and scaled to
    all of them.  Returns (seconds for all couples, negatives of the sampled couples)."""
    import oracle
    t0 = time.perf_counter()
    S = oracle.mining.all_pairs_similarities(E)
    t_mm = time.perf_counter() - t0
    couples = list(zip(anchors[:n_couples].tolist(), positives[:n_couples].tolist()))
    t0 = time.perf_counter()
    neg = oracle.select_negatives(S, lab, couples, semi)
    t_sel = time.perf_counter() - t0
    return t_mm + t_sel * (anchors.numel() / float(n_couples)), t_mm, neg


def side_mining(dev, pk, cpu_baseline=True):
    """configs[2]: all-pairs similarities + (semi-)hard negative of every anchor, 16384 x 2048."""
    from instance_search_b200 import mining
    g = torch.Generator(device=dev).manual_seed(1234 + 3)
    N, D, per = 16384, 2048, 16
    lab = torch.arange(N, device=dev) // per
    E = torch.randn(N // per, D, device=dev, generator=g)[lab] + 0.5 * torch.randn(N, D, device=dev, generator=g)
    E = E / E.norm(dim=1, keepdim=True)
    anchors = torch.arange(N, device=dev)
    positives = (anchors // per) * per + (anchors % per + 1) % per
    idx = mining.MiningIndex(E, lab.int())
    out = {"workload": "negative mining, 16384 x 2048-d descriptors, 16 per label, one couple per anchor "
                       "(BASELINE configs[2])"}
    flops = 2.0 * N * N * D          # 2 N^2 D, whatever the screen issues to get there
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    E_host, lab_host = (E.cpu(), lab.cpu()) if cpu_baseline else (None, None)
    for semi in (True, False):
        ms = _median_ms(lambda: idx.select_negatives(anchors, positives, semi))
        tf = flops / (ms * 1e-3) / 1e12
        rec = {"ms": ms, "anchors_per_s": N / (ms * 1e-3),
               "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                            "algorithmic_flops": flops},
               "bruteforce_rows": int(idx.last_bruteforce), "screen_terms": idx.terms}
        if cpu_baseline:
            torch.set_num_threads(os.cpu_count() or 1)
            nc = 1024
            dt, t_mm, want = cpu_mining(E_host, lab_host, anchors.cpu(), positives.cpu(), semi, nc)
            got = idx.select_negatives(anchors, positives, semi)[0][:nc].cpu()
            rec["cpu_baseline"] = {"value": N / dt, "unit": "anchors/s", "cores": torch.get_num_threads(),
                                   "kind": "port", "sample": "mm(E, E.t()) in full (%.1f s) + masked arg-max of the "
                                   "first %d couples scaled to %d (oracle, torch fp32)" % (t_mm, nc, N)}
            rec["parity"] = {"couples": nc, "identical_negatives": int((got == want).sum())}
        out["semi_hard" if semi else "hard"] = rec
    return out


def side_e2e(a, dev, rank, world, cpu_baseline=True):
    """BASELINE configs[4]: ResNet-152 trunk (PyTorch, random init, 448-px input -> 14 x 14 maps) ->
    region descriptors (fused CUDA head, D = 512, k = 6) -> top-100 over a synthetic 10M x 512-d
    database, row-sharded over the N ranks (images data-parallel, descriptors all-gathered, candidate
    exchange over NCCL).  Collective: every rank calls it.  Rank 0 also checks the returned top-100 of
    a query sample against the CPU oracle over the full database."""
    import torch.distributed as dist
    import torchvision
    from instance_search_b200 import regions
    from instance_search_b200.model.siamese import RegionDescriptorNet
    from instance_search_b200.search import ShardedIndex, shard_bounds
    from instance_search_b200.sharding import all_gather_rows
    n_db, dim, k, per_gpu, batch, px = a.e2e_db_rows, 512, 100, 128, 32, 448
    torch.manual_seed(0)
    trunk = torchvision.models.resnet152(weights=None, num_classes=464)
    net = RegionDescriptorNet(trunk, 6, dim, (7, 7)).to(dev).eval()
    lo, hi = shard_bounds(n_db, world)[rank]
    index = ShardedIndex(make_rows_slice(n_db, dim, 1234 + 5, dev, lo, hi), n_db, rank, world)
    g = torch.Generator().manual_seed(100 + rank)
    mean = torch.tensor([0.36, 0.30, 0.28]).view(1, 3, 1, 1)
    std = torch.tensor([0.21, 0.20, 0.20]).view(1, 3, 1, 1)
    images = ((torch.rand(per_gpu, 3, px, px, generator=g) - mean) / std).pin_memory()
    n_img = per_gpu * world

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def one_pass():
        marks, descs = [], []
        e = [ev() for _ in range(3)]
        e[0].record()
        with torch.no_grad():
            for s in range(0, per_gpu, batch):
                x = images[s:s + batch].to(dev, non_blocking=True)
                m0, m1, m2 = ev(), ev(), ev()
                m0.record()
                fmap = net.features(x)                    # PyTorch trunk (north_star: stays in PyTorch)
                m1.record()
                descs.append(regions.region_descriptors(fmap, net._head(), net.k, net.feature_size2d,
                                                        want_cls_out=False)[0])
                m2.record()
                marks.append((m0, m1, m2))
        q = all_gather_rows(torch.cat(descs), n_img, rank, world)       # [n_img, 512] on every rank
        e[1].record()
        scores, idx = index.search(q, k)
        e[2].record()
        return marks, e, q, scores, idx

    times = []
    for it in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        marks, e, q, scores, idx = one_pass()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if it >= 1:
            times.append((sum(m0.elapsed_time(m1) for m0, m1, _ in marks),
                          sum(m1.elapsed_time(m2) for _, m1, m2 in marks), e[1].elapsed_time(e[2]),
                          e[0].elapsed_time(e[2])))
    best = min(times, key=lambda t: t[3])
    ms = torch.tensor(best, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    tr, hd, se, tot = [float(v) for v in ms.tolist()]
    out = None
    if rank == 0:
        out = {"workload": "end-to-end instance search (BASELINE configs[4]): ResNet-152 trunk (PyTorch, %d px) -> region "
                           "descriptors (D=%d, k=6) -> top-%d over %d x %d-d database, %d GPU(s), %d images per GPU"
                           % (px, dim, k, n_db, dim, world, per_gpu),
               "n_gpus": world, "images": n_img, "ms": {"trunk": tr, "region_head": hd, "search": se, "total": tot},
               "images_per_s_end_to_end": n_img / (tot * 1e-3), "search_queries_per_s": n_img / (se * 1e-3),
               "region_head_images_per_s": n_img / (hd * 1e-3), "parity": None}
        if cpu_baseline:
            nq = min(32, n_img)
            db_host = make_rows_host(n_db, dim, 1234 + 5, dev)
            qh = q[:nq].cpu()
            cs, ci = cpu_topk(qh, db_host, k)
            out["parity"] = topk_parity(qh, db_host, k, scores[:nq], idx[:nq], cs, ci)
            del db_host
    del index, net
    torch.cuda.empty_cache()
    return out


def make_rows_host(n, d, seed, device, chunk=131072):
    """make_rows(n, d, seed) drawn chunk by chunk on `device` (the same global stream every shard
    is cut from) and moved to host memory: the oracle's copy of a database that is sharded over
    several GPUs."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((n, d), dtype=torch.float32)
    for s in range(0, n, chunk):
        x = torch.randn((min(chunk, n - s), d), generator=g, device=device)
        out[s:s + chunk] = (x / x.norm(dim=1, keepdim=True)).cpu()
    return out


def make_rows_slice(n, d, seed, device, lo, hi, chunk=131072):
    """Rows [lo, hi) of make_rows(n, d, seed): the global stream is drawn chunk by
    chunk on every rank, only the local rows are kept."""
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((hi - lo, d), dtype=torch.float32, device=device)
    for s in range(0, n, chunk):
        e = min(s + chunk, n)
        x = torch.randn((e - s, d), generator=g, device=device)
        a, b = max(s, lo), min(e, hi)
        if a < b:
            rows = x[a - s:b - s]
            out[a - lo:b - lo] = rows / rows.norm(dim=1, keepdim=True)
        if e >= hi:
            break
    return out


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    if world != a.gpus and world == 1 and a.gpus > 1:
        # launched without torchrun: re-exec under it
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               "--nproc-per-node", str(a.gpus), "--master-addr", "127.0.0.1", "--master-port", "29517",
               os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_b200(a, rank, world, local_rank)


if __name__ == "__main__":
    main()
