#!/usr/bin/env python
"""Benchmark of the interaction-scoring hot path (BASELINE.json metric: query-doc pairs scored / s,
ColBERT max-sim, dim=128, Lq=32, Ld=180, 64 queries x 1000 docs per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload colbert]

N > 1 is launched by torchrun (one rank per GPU, NCCL).  Prints ONE JSON line on rank 0.

A "step" = one pass of the hot path over one synthetic batch (workload below).  `value` is whole-job
pairs/s with inputs resident in HBM; `e2e` is the same metric through the host-buffer C-ABI call
(pinned host inputs -> H2D -> kernel -> D2H scores inside the timed region); `roofline` is the max-sim
kernel's algorithmic bytes / its CUDA-event duration against the measured HBM peak; `cpu_baseline` is
the CPU oracle (port of the reference's PyTorch path, timed on this host's cores).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

# BASELINE config 3
N_QUERIES, DOCS_PER_QUERY, LQ, LD, DIM = 64, 1000, 32, 180, 128
SEED = 1237
TOPK = 100
# doc tile Ld*dim*2 + 4 B length + 4 B score + query tile amortised over 1000 docs
ALG_BYTES_PER_PAIR = LD * DIM * 2 + 4 + 4 + (LQ * DIM * 2) // DOCS_PER_QUERY
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled every ~5 ms through NVML while the timed regions run (nvidia-smi
    -lms 200 is too coarse: the 20-step timed region of this kernel lasts ~10 ms)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = threading.Event()
        self.thread = None
        self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except Exception:
                    pass
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.smax = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nv
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag.is_set():
            try:
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = get_reasons(self.h)
                self.samples.append((time.perf_counter(), clk, rs))
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                break
            time.sleep(0.005)

    def stop(self, windows):
        """windows: list of (t0, t1) perf_counter intervals that were timed regions."""
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
        self.stop_flag.set()
        self.thread.join(timeout=2)
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
        inside = [s for s in self.samples if any(a <= s[0] <= b for a, b in windows)]
        used = inside if inside else self.samples
        clocks = [c for _, c, _ in used]
        reasons = set()
        for _, _, r in used:
            for n, bit in names.items():
                if r & bit:
                    reasons.add(n)
        return {"sm_mhz": statistics.median(clocks) if clocks else None, "sm_max_mhz": self.smax,
                "samples_in_timed_regions": len(inside), "samples_total": len(self.samples),
                "reasons": sorted(reasons), "how": "NVML poll every 5 ms; timed regions = value loop + e2e loop"}


def make_inputs(shard: int):
    from oracle import interaction_oracle as O  # input generator only (shared with the tests)
    return O.synth_colbert_inputs(N_QUERIES, DOCS_PER_QUERY, LQ, LD, DIM, seed=SEED + shard)


def cpu_oracle_step(q32, d32, qm, dm):
    from oracle import interaction_oracle as O
    with torch.no_grad():
        return O.maxsim_one_query_many_docs(q32, d32, qm, dm, DOCS_PER_QUERY)


def time_cpu_baseline(q, d, qm, dm, budget_s: float = 12.0):
    """Reference arithmetic (fp32 upcast of the fp16 storage, dense_retrieval.py:406) on the host cores."""
    q32, d32 = q.float(), d.float()
    cpu_oracle_step(q32[:4], d32[:4 * DOCS_PER_QUERY], qm[:4], dm[:4 * DOCS_PER_QUERY])  # warm-up
    reps, t_total = 0, 0.0
    while t_total < budget_s and reps < 50:
        t0 = time.perf_counter()
        cpu_oracle_step(q32, d32, qm, dm)
        t_total += time.perf_counter() - t0
        reps += 1
    pairs = N_QUERIES * DOCS_PER_QUERY * reps
    return {"value": pairs / t_total, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{reps} x full workload ({N_QUERIES} queries x {DOCS_PER_QUERY} docs), torch CPU fp32, "
                      f"{torch.get_num_threads()} threads of {os.cpu_count()} logical cores"}


def config_dict(n_gpus):
    return {"workload": "colbert_maxsim", "queries_per_gpu": N_QUERIES, "docs_per_query": DOCS_PER_QUERY,
            "Lq": LQ, "Ld": LD, "dim": DIM, "storage_dtype": "float16", "mask_dtype": "bool",
            "pairs_per_step": N_QUERIES * DOCS_PER_QUERY * n_gpus,
            "sharding": "documents sharded over ranks; per-query top-%d all-gather + merge when N>1" % TOPK,
            "l2_policy": "inputs larger than L2 (2.95 GB of documents per GPU per step vs 126 MB L2)"}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU PyTorch path (oracle port) on this host, rank 0 only."""
    if rank != 0:
        return
    q, d, qm, dm = make_inputs(0)
    q32, d32 = q.float(), d.float()
    for _ in range(args.warmup):
        cpu_oracle_step(q32, d32, qm, dm)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_oracle_step(q32, d32, qm, dm)
    dt = time.perf_counter() - t0
    pairs = N_QUERIES * DOCS_PER_QUERY
    v = pairs * args.steps / dt
    cfg = config_dict(1)
    cfg["pairs_per_step"] = pairs
    line = {"impl": "reference", "metric": "query-doc pairs scored/sec (ColBERT max-sim d=128)", "value": v,
            "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"each step = full workload {N_QUERIES}x{DOCS_PER_QUERY} pairs on "
                                       f"{torch.get_num_threads()} torch threads ({os.cpu_count()} logical cores)"},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed end-to-end steps (default: min(steps, 5))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from matchmaker_b200 import interaction, sharding

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    q, d, qm, dm = make_inputs(rank)
    qm_b, dm_b = qm.bool(), dm.bool()
    cq, cd, cqm, cdm = q.to(dev), d.to(dev), qm_b.to(dev), dm_b.to(dev)
    doc_id_base = rank * N_QUERIES * DOCS_PER_QUERY

    def step():
        s = interaction.maxsim(cq, cd, cqm, cdm, docs_per_query=DOCS_PER_QUERY, impl="tcgen05")
        if world > 1:
            return sharding.topk_all_gather_merge(s.view(N_QUERIES, DOCS_PER_QUERY), TOPK, doc_id_base)
        return s

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()

    # ---- timed region: `value` (inputs resident in HBM) -----------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    kern_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_val0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        kern_ev[i][0].record()
        s = interaction.maxsim(cq, cd, cqm, cdm, docs_per_query=DOCS_PER_QUERY, impl="tcgen05")
        kern_ev[i][1].record()
        if world > 1:
            sharding.topk_all_gather_merge(s.view(N_QUERIES, DOCS_PER_QUERY), TOPK, doc_id_base)
    e1.record()
    sync_all()
    t_val1 = time.perf_counter()
    ms_total = e0.elapsed_time(e1)
    kern_ms = statistics.mean(a.elapsed_time(b) for a, b in kern_ev)
    t = torch.tensor([ms_total, kern_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, kern_ms = t.tolist()
    ms_per_step = ms_total / args.steps
    pairs_per_step = N_QUERIES * DOCS_PER_QUERY * world
    value = pairs_per_step / (ms_per_step * 1e-3)

    # ---- e2e: host-buffer C-ABI call, H2D + kernel + D2H inside the timed region ------------------
    hq, hd, hqm, hdm = q.pin_memory(), d.pin_memory(), qm_b.pin_memory(), dm_b.pin_memory()
    e2e_steps = args.e2e_steps or min(args.steps, 5)
    for _ in range(2):
        interaction.maxsim_host(hq, hd, hqm, hdm, docs_per_query=DOCS_PER_QUERY, device=dev)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = interaction.maxsim_host(hq, hd, hqm, hdm, docs_per_query=DOCS_PER_QUERY, device=dev)  # synchronous
    torch.cuda.synchronize()
    t_e2e1 = time.perf_counter()
    clocks = sampler.stop([(t_val0, t_val1), (t0, t_e2e1)]) if rank == 0 else None
    e2e_s = (t_e2e1 - t0) / e2e_steps
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = te.item()
    h2d = sum(x.numel() * x.element_size() for x in (hq, hd, hqm, hdm))
    d2h = out.numel() * out.element_size()

    if rank == 0:
        peak, peak_src = _peaks()
        alg_bytes = ALG_BYTES_PER_PAIR * N_QUERIES * DOCS_PER_QUERY
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": "query-doc pairs scored/sec (ColBERT max-sim d=128)", "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": config_dict(world),
            "clocks": clocks,
            "e2e": {"value": pairs_per_step / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3,
                    "note": "mmb200_maxsim_fwd_host: pinned host q/d/masks -> chunked H2D overlapped with the "
                            "kernel -> D2H scores; PCIe-bound"},
            "gpu_launches": args.steps,
            "roofline": {"bound": "hbm", "kernel": "maxsim_qm_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "kernel_ms": kern_ms, "algorithmic_bytes_per_launch": alg_bytes,
                         "algorithmic_bytes_per_pair": ALG_BYTES_PER_PAIR},
        }
        prof = os.path.join(ROOT, "profiles", "maxsim_traffic.json")
        if os.path.isfile(prof):
            try:
                line["roofline"]["traffic"] = json.load(open(prof))["dram_bytes_per_launch"]
            except Exception:
                pass
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = time_cpu_baseline(q, d, qm, dm)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
