#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched FM-index search path (BASELINE.json configs[1]):
parallel_count of 10 M synthetic 150-bp reads against a 3.1 Gbp synthetic DNA text, k = 13 seed
table, SA ratio 8.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (oracle port)

One process per GPU (torchrun for N > 1); the index is replicated per GPU, every rank searches its
own 10 M-read batch (weak scaling), no collective on the data path.  A "step" is one pass of the hot
path over the rank's batch.  Rank 0 prints ONE JSON line.
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

TEXT_SEED = 3
QUERY_SEED = 4
ALG_BYTES_PER_BLOCK = 104   # 3 x 32 B planes + one 8-B milestone per rank-block access


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="awry_b200", choices=["awry_b200", "reference"])
    ap.add_argument("--text-len", type=int, default=3_100_000_000)
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--kmer", type=int, default=13)
    ap.add_argument("--sa-ratio", type=int, default=8)
    ap.add_argument("--cpu-sample", type=int, default=4_000_000, help="reads timed on the CPU oracle (rank 0, N=1)")
    ap.add_argument("--ref-reads-per-step", type=int, default=1_000_000)
    ap.add_argument("--locate-reads", type=int, default=1_000_000, help="cfg3: queries of the locate leg")
    ap.add_argument("--locate-len", type=int, default=50)
    ap.add_argument("--no-locate", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  The sampler is
    started ahead of the region and `wait_ready()` returns once its first line has arrived; only samples whose
    timestamp falls inside [mark_begin(), mark_end()] are reported (all of them if none does)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def wait_ready(self, timeout=3.0):
        t = time.time()
        while self.proc and not self.lines and time.time() - t < timeout:
            time.sleep(0.01)

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def parse(rows):
            sm, mx, reasons = [], [], set()
            for _, ln in rows:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 10:
                    continue
                try:
                    sm.append(float(f[2]))
                    mx.append(float(f[3]))
                except ValueError:
                    continue
                for name, v in zip(names, f[6:10]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            return sm, mx, reasons
        # a line is printed right after its sample is taken: arrival time ~ sample time
        inside = [r for r in self.lines if self.t0 is not None and self.t0 - 0.01 <= r[0] <= (self.t1 or 1e18) + 0.06]
        sm, mx, reasons = parse(inside if inside else self.lines)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "samples_in_timed_region": len(inside), "reasons": sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, stream copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def traffic_from_profiles():
    """dram bytes per launch of the search kernel from the committed ncu capture, if any"""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "search_kernel_traffic.json")))
        return d
    except Exception:
        return None


def workload_name(a):
    return (f"parallel_count {a.reads} x {a.read_len}-bp reads vs {a.text_len}-bp synthetic DNA text, "
            f"k={a.kmer} seed table, SA ratio {a.sa_ratio}")


def build_host_index(a, device):
    """reference-layout index arrays in host memory (fixture; FmIndex::new stand-in)"""
    import torch
    if torch.cuda.is_available():
        from fixtures import pyfixture_gpu as fxg
        parts, phases = fxg.build_parts(0, a.text_len, TEXT_SEED, ratio=a.sa_ratio, kmer_len=a.kmer, device=device)
        return parts, phases
    from fixtures import pyfixture as fx
    text = fx.gen_text(0, a.text_len, TEXT_SEED)
    return fx.build_parts(text, 0, ratio=a.sa_ratio, kmer_len=a.kmer), {"total": None}


def host_reads(a, nq, qseed):
    """ASCII reads of the synthetic text regenerated from its seed (no stored text needed)"""
    import torch
    if torch.cuda.is_available():
        from fixtures import pyfixture_gpu as fxg
        d = torch.empty(nq * a.read_len, dtype=torch.uint8, device="cuda")
        fxg.gen_queries_device(0, a.text_len, TEXT_SEED, nq, a.read_len, qseed, d.data_ptr())
        torch.cuda.synchronize()
        return d
    raise RuntimeError("no GPU")


def run_reference(a):
    """--impl reference: the reference's CPU search path (C port in oracle/, all host threads)"""
    rank, local_rank, world = dist_env()
    if rank != 0:
        return 0
    from oracle import pyoracle as po
    import torch
    if torch.cuda.is_available():
        torch.cuda.set_device(0)
    reduced = not torch.cuda.is_available()
    if reduced:
        a.text_len = min(a.text_len, 20_000_000)
    parts, _ = build_host_index(a, 0)
    orc = po.OracleIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks,
                                    parts.prefix_sums, parts.sa_words)
    nq = a.ref_reads_per_step
    if reduced:
        from fixtures import pyfixture as fx
        text = fx.gen_text(0, a.text_len, TEXT_SEED)
        qb, _, _ = fx.gen_substring_queries(text, nq, a.read_len, QUERY_SEED)
    else:
        qb = host_reads(a, nq, QUERY_SEED).cpu().numpy()
    qo = np.arange(nq + 1, dtype=np.uint64) * np.uint64(a.read_len)
    cores = po.lib().awo_hw_threads()
    for _ in range(a.warmup):
        orc.count_batch(qb, qo)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        counts, st = orc.count_batch(qb, qo)
    dt = time.perf_counter() - t0
    value = nq * a.steps / dt
    sample = (f"{nq} reads per step (bounded sample of the {a.reads}-read batch), plain backward search as the "
              f"reference executes it (no table use, kmer_lookup_table.rs:90-110), {cores} pthreads with dynamic "
              f"chunking standing in for rayon")
    line = {
        "impl": "reference", "metric": "count queries/s", "value": value, "unit": "reads/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(a), "reduced_text": reduced, "reads_per_step": nq},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    a = parse_args()
    if a.impl == "reference":
        return run_reference(a)
    rank, local_rank, world = dist_env()
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        print("bench.py: no CUDA device; the awry_b200 search path has no CPU fallback", file=sys.stderr)
        return 2
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from awry_b200 import FmIndex, fm_index as f
    from fixtures import pyfixture_gpu as fxg

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- setup (untimed): index fixture, device replica, reads in HBM and in pinned host memory
    t_setup = time.time()
    parts, phases = build_host_index(a, local_rank)
    ix = FmIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks,
                            parts.prefix_sums, parts.sa_words, devices=[local_rank])
    nq, L = a.reads, a.read_len
    qseed = QUERY_SEED + 1000 * rank           # every rank searches its own batch
    d_q = torch.empty(nq * L, dtype=torch.uint8, device="cuda")
    fxg.gen_queries_device(0, a.text_len, TEXT_SEED, nq, L, qseed, d_q.data_ptr())
    d_off = torch.arange(0, nq + 1, dtype=torch.int64, device="cuda") * L
    d_cnt = torch.zeros(nq, dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()
    setup_s = time.time() - t_setup

    def step_device():
        ix.count_device(d_q.data_ptr(), d_off.data_ptr(), nq, d_cnt.data_ptr(), stream)

    # ---- kernel-only figure: inputs resident in HBM, CUDA events on the launching stream
    for _ in range(max(a.warmup, 3)):
        step_device()
    barrier()
    f.profile_enable(True)
    f.profile_reset()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_ready()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record()
    for _ in range(a.steps):
        step_device()
    e1.record()
    barrier()
    sampler.mark_end()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    prof = f.profile_get()
    f.profile_enable(False)
    ix.device_check(stream)
    t_ms = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * nq * a.steps / (ms_max * 1e-3)

    # ---- end-to-end figure: C-ABI call with pinned HOST buffers, H2D + D2H inside the timed region
    e2e = None
    if not a.no_e2e:
        h_q = torch.empty(nq * L, dtype=torch.uint8, pin_memory=True)
        h_q.copy_(d_q)
        h_off = torch.empty(nq + 1, dtype=torch.int64, pin_memory=True)
        h_off.copy_(d_off)
        h_cnt = torch.empty(nq, dtype=torch.int64, pin_memory=True)
        torch.cuda.synchronize()
        qb = h_q.numpy()
        qo = h_off.numpy().view(np.uint64)
        out = h_cnt.numpy().view(np.uint64)
        for _ in range(2):
            ix.count_packed(qb, qo, out=out)
        f.profile_reset()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            ix.count_packed(qb, qo, out=out)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        p2 = f.profile_get()
        t_e = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        e2e = {"value": world * nq * a.steps / float(t_e.item()), "unit": "reads/s",
               "h2d_bytes_per_step": p2["h2d_bytes"] // a.steps, "d2h_bytes_per_step": p2["d2h_bytes"] // a.steps,
               "ms_per_step": float(t_e.item()) / a.steps * 1e3, "host_buffers": "pinned",
               "host_pack": ("bases packed to 2 bits by the host thread pool before the copy"
                             if p2["h2d_bytes"] // a.steps < nq * L else "ASCII bytes copied as they are")}
        assert np.array_equal(out, d_cnt.cpu().numpy().view(np.uint64)), "e2e and device-resident counts differ"

    # ---- secondary run of cfg2: 10 % of the reads carry one random substitution, so their
    # search ends early with an empty interval; same timing rules as the headline figure
    mutated = None
    if not a.no_locate:
        d_qm = torch.empty(nq * L, dtype=torch.uint8, device="cuda")
        fxg.gen_queries_device(0, a.text_len, TEXT_SEED, nq, L, qseed, d_qm.data_ptr(), mut_ppm=100_000)
        d_cm = torch.zeros(nq, dtype=torch.int64, device="cuda")
        for _ in range(3):
            ix.count_device(d_qm.data_ptr(), d_off.data_ptr(), nq, d_cm.data_ptr(), stream)
        barrier()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        for _ in range(a.steps):
            ix.count_device(d_qm.data_ptr(), d_off.data_ptr(), nq, d_cm.data_ptr(), stream)
        m1.record()
        barrier()
        t_m = torch.tensor([m0.elapsed_time(m1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_m, op=dist.ReduceOp.MAX)
        zero = int((d_cm == 0).sum())
        mutated = {"workload": "same batch, 10 % of the reads with one substitution (early exit)",
                   "reads_per_s": world * nq * a.steps / (float(t_m.item()) * 1e-3),
                   "ms_per_step": float(t_m.item()) / a.steps, "reads_without_a_hit": zero}
        del d_qm, d_cm

    # ---- secondary metric (BASELINE cfg3): parallel_locate of 1 M x 50-bp queries, same index
    locate = None
    if not a.no_locate:
        nl, ll = a.locate_reads, a.locate_len
        d_lq = torch.empty(nl * ll, dtype=torch.uint8, device="cuda")
        fxg.gen_queries_device(0, a.text_len, TEXT_SEED, nl, ll, QUERY_SEED + 1 + 1000 * rank, d_lq.data_ptr())
        d_loff = torch.arange(0, nl + 1, dtype=torch.int64, device="cuda") * ll
        d_hoff = torch.zeros(nl + 1, dtype=torch.int64, device="cuda")
        def locate_leg(variant):
            """device-resident two-pass locate, CUDA events; variant 0 = default pass 2, 1 = LF-walk"""
            f.set_locate_variant(variant)
            try:
                nh = 0
                for _ in range(2):
                    ptr, nh = ix.locate_device(d_lq.data_ptr(), d_loff.data_ptr(), nl, d_hoff.data_ptr(), stream=stream)
                    ix.device_free(ptr)
                f.profile_enable(True)
                f.profile_reset()
                barrier()
                l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                l0.record()
                for _ in range(a.steps):
                    ptr, nh = ix.locate_device(d_lq.data_ptr(), d_loff.data_ptr(), nl, d_hoff.data_ptr(), stream=stream)
                    ix.device_free(ptr)
                l1.record()
                barrier()
                prof_l = f.profile_get()
                f.profile_enable(False)
            finally:
                f.set_locate_variant(0)
            t = torch.tensor([l0.elapsed_time(l1)], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item()) / a.steps, prof_l, nh

        walk_lms, walk_lp, n_hits = locate_leg(1)
        lms, lp, n_hits2 = locate_leg(0)
        assert n_hits == n_hits2
        unsampled = ix.device_bytes()["full_sa"] > 0
        # end to end through the C ABI: pinned host queries in, CSR offsets + hits out to the host
        hl_q = torch.empty(nl * ll, dtype=torch.uint8, pin_memory=True)
        hl_q.copy_(d_lq)
        hl_off = torch.empty(nl + 1, dtype=torch.int64, pin_memory=True)
        hl_off.copy_(d_loff)
        torch.cuda.synchronize()
        lqb, lqo = hl_q.numpy(), hl_off.numpy().view(np.uint64)
        hoff_h = torch.zeros(nl + 1, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
        hits_h = torch.zeros((n_hits + 1024, 2), dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
        ix.locate_packed_into(lqb, lqo, hoff_h, hits_h)
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            n_e2e = ix.locate_packed_into(lqb, lqo, hoff_h, hits_h)
        dt = (time.perf_counter() - t0) / a.steps
        t_le = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_le, op=dist.ReduceOp.MAX)
        locate = {"workload": f"parallel_locate {nl} x {ll}-bp queries, SA ratio {a.sa_ratio} (cfg3)",
                  "pass2": "gather from the unsampled suffix array (rebuilt on the device at load time)" if unsampled
                           else "LF-walk to the sampled rows",
                  "hits_per_query": n_hits / nl, "hits_per_s": world * n_hits / (lms * 1e-3),
                  "queries_per_s": world * nl / (lms * 1e-3), "ms_per_step": lms,
                  "pass2_kernel_ms": lp["walk_ms"] / max(1, lp["walk_launches"]),
                  "search_kernel_ms": lp["search_ms"] / max(1, lp["search_launches"]),
                  "e2e_hits_per_s": world * n_e2e / float(t_le.item()),
                  "e2e_ms_per_step": float(t_le.item()) * 1e3,
                  "lf_walk_variant": {"hits_per_s": world * n_hits / (walk_lms * 1e-3), "ms_per_step": walk_lms,
                                      "walk_kernel_ms": walk_lp["walk_ms"] / max(1, walk_lp["walk_launches"])}}
        del d_lq, d_loff, d_hoff

    # ---- CPU baseline + parity + exact algorithmic work (rank 0, N = 1 only)
    cpu_baseline, parity, touches_per_read = None, None, None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle import pyoracle as po
        orc = po.OracleIndex.from_parts(parts.alphabet, parts.ratio, parts.bwt_len, parts.kmer_len, parts.blocks,
                                        parts.prefix_sums, parts.sa_words)
        ns = min(a.cpu_sample, nq)
        sb = d_q[: ns * L].cpu().numpy()
        so = np.arange(ns + 1, dtype=np.uint64) * np.uint64(L)
        cores = po.lib().awo_hw_threads()
        orc.count_batch(sb[: 100_000 * L], so[:100_001])     # warm-up
        t0 = time.perf_counter()
        want, st = orc.count_batch(sb, so)
        dt = time.perf_counter() - t0
        parity = bool(np.array_equal(want, d_cnt[:ns].cpu().numpy().view(np.uint64)))
        touches_per_read = st["seeded_touches"] / ns
        cpu_baseline = {"value": ns / dt, "unit": "reads/s", "cores": cores, "kind": "port",
                        "sample": f"first {ns} reads of the batch ({dt:.1f} s), C port of the reference search path "
                                  f"in the reference block layout, {cores} pthreads (dynamic chunks) for rayon"}
    if touches_per_read is None:
        # closed form of SURVEY.md 8(d): L-k seeded steps, ~1.006 distinct blocks per step
        touches_per_read = (L - a.kmer) * 1.006

    # ---- roofline of the dominant kernel (backward search), measured live
    peak, peak_src = measured_peak()
    alg_bytes_per_launch = nq * (L + 8 + 16 + ALG_BYTES_PER_BLOCK * touches_per_read)
    search_ms = prof["search_ms"] / max(1, prof["search_launches"])
    achieved = alg_bytes_per_launch / (search_ms * 1e-3) / 1e9
    tr = traffic_from_profiles()
    gather = None
    if rank == 0:
        try:   # the "random-access HBM roofline" of the north star: independent random 128-B reads
            # 4 lanes x LDG.256 per read, 4 in flight per lane group, 16 waves of blocks so the hardware
            # balances the SMs (a static split stops at 37.8 G reads/s: it ends with the slowest SM)
            g_reads, g_gbs = f.bench_random_gather(local_rank, 4 << 30, 128, 3164, 400_000_000, 2)
            accesses = nq * (((L - a.kmer) + 1) // 2 + 1)          # pair steps (+ odd tail) + seed lookup
            gather = {"granule_bytes": 128, "reads_per_s": g_reads, "gb_per_s": g_gbs,
                      "kernel_block_reads_per_s": accesses / (search_ms * 1e-3),
                      "frac_of_random_gather": accesses / (search_ms * 1e-3) / g_reads}
        except Exception as e:  # noqa: BLE001
            gather = {"error": str(e)}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": tr["dram_bytes_per_launch"] if tr else None, "peak_source": peak_src,
                "kernel": "search_dna_pair_kernel", "kernel_ms": search_ms,
                "kernel_share_of_step": prof["search_ms"] / ms_total,
                "algorithmic_bytes_per_launch": alg_bytes_per_launch,
                "lf_steps_per_s": nq * (L - a.kmer) / (search_ms * 1e-3), "random_gather_roofline": gather,
                "note": "achieved counts the reference algorithm's bytes (104 B per LF step); the pair index does two "
                        "LF steps per 128-B block read, so achieved can exceed both `traffic` and the stream peak"}

    if rank == 0:
        line = {
            "metric": "count queries/s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_max / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": workload_name(a), "reads_per_gpu_per_step": nq,
                       "l2": "inputs larger than L2 (1.5 GB reads, 1.55 GB rank blocks, 0.5 GB seed table per step)",
                       "index": "replicated per GPU; ranks search disjoint batches; no collective on the data path",
                       "setup_s": round(setup_s, 1), "fixture_build_s": phases.get("total")},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
            "gpu_launches": int(prof["launches"]), "parity_vs_oracle_on_sample": parity, "locate": locate,
            "count_with_mismatches": mutated,
        }
        print(json.dumps(line), flush=True)
    ix.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
