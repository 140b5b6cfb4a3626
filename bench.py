#!/usr/bin/env python
"""Benchmark of the GeneralGNN hot path (BASELINE.json metric: train graphs/sec and
GCN-layer edges/sec at 1/2/4/8 B200; SpMM HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one full training step on one disjoint batch of 1024 synthetic E. coli-shaped
graphs per GPU (~500 nodes, stored degree ~12, 32-d features; default GeneralGNN: hidden 256,
4 GeneralConv layers, sum aggregation, sum pool, 2 classes): device batching (K0) -> forward
-> categorical cross-entropy -> backward -> [NCCL gradient all-reduce] -> fused SGD.  That is
the per-GPU step of BASELINE.json configs[2] (same shape as configs[1], which is its forward
half; the forward-only rate is reported beside it as `fwd_graphs_per_sec`).  Graphs shard by
graph across ranks (weak scaling: 1024 graphs per GPU per step).

`value`  : graphs/s, dataset resident in HBM, timed with CUDA events, max over ranks.
`e2e`    : the same step through the public loader with the dataset in PINNED HOST memory:
           every step uploads its graphs (H2D) and reads the loss back (D2H) inside the timing.
`roofline`: the aggregation kernel (SpMM fwd), algorithmic bytes / CUDA-event time vs the
           measured HBM peak in MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the CPU restatement of the reference's op sequence
           (oracle O2, PyTorch-CPU fp32; kind "port" — TensorFlow/Spektral cannot be installed
           here) on the host cores, bounded sample.
"""
from __future__ import annotations

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

B_GRAPHS = 1024
N_MEAN, DEG, N_FEAT, HIDDEN, LAYERS, CLASSES = 500, 12, 32, 256, 4, 2
WORKLOAD = ("cfg3 per-GPU train step (= cfg2 shape): GeneralGNN hidden 256 x 4 GeneralConv, sum agg, sum pool, "
            "1024 synthetic E.coli-shaped graphs/GPU/step (~500 nodes, deg ~12, 32-d features)")


def _peaks():
    """(HBM GB/s, sustained bf16 TFLOP/s, source) - measured by the driver on this pool, else the recipe's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1400.0))), "measured"
    except Exception:
        return 6650.0, 1400.0, "fallback"


def _ncu_traffic():
    """DRAM bytes per launch of the roofline kernel from the committed ncu --set full capture (or None)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_spmm_rb4_ncu.json")) as f:
            d = json.load(f)
        return float(d["dram_bytes_read"]) + float(d["dram_bytes_write"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.t0 = self.t1 = None

    def mark_begin(self):
        """Start of the timed region: the sampler itself is started earlier (before the warm-up steps), because
        nvidia-smi's start-up (NVML initialisation) stalls kernel launches for tens of milliseconds."""
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def wait_ready(self, timeout=3.0):
        """Block until nvidia-smi has delivered its first line (its initialisation is over)."""
        t = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t < timeout:
            time.sleep(0.01)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = [ln for t, ln in self.lines if (self.t0 is None or t >= self.t0) and (self.t1 is None or t <= self.t1 + 0.03)]
        if not lines:                                     # a timed region shorter than one sampling period
            lines = [ln for t, ln in self.lines if self.t0 is None or t >= self.t0] or [ln for _, ln in self.lines[-1:]]
        for ln in lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_rate(n_graphs, steps, warmup, train=True):
    """graphs/s of the CPU restatement (oracle O2) on `n_graphs` graphs of the bench workload."""
    import torch
    import gcn_string_b200 as g
    from gcn_string_b200 import synthetic
    from oracle import model_ref_torch as O2
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    ds = synthetic.make_dataset(n_graphs, seed=0, n_mean=N_MEAN, deg=DEG, n_feat=N_FEAT)
    graphs = [ds.graph(k) for k in range(n_graphs)]
    cfg = g.GNNConfig(in_features=N_FEAT, output=CLASSES, activation="softmax", hidden=HIDDEN, message_passing=LAYERS)
    w, s = g.init_params(cfg, seed=0)
    sec, threads = O2.time_reference_path(cfg, g.block_specs(cfg), w, s, graphs, n_steps=steps, warmup=warmup,
                                          train=train, lr=0.0002, threads=cores)
    return n_graphs / sec, threads, sec, int(ds.n_edges.sum())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n = 64
    rate, threads, sec, nnz = cpu_reference_rate(n, max(1, args.steps), max(1, args.warmup))
    sample = (f"{n} graphs/step of the same workload (host scipy collate + fwd + bwd + SGD, PyTorch-CPU fp32 restatement "
              f"of the reference op sequence; TensorFlow/Spektral not installable here), median of {max(1, args.steps)} steps")
    line = {"impl": "reference", "metric": "train_graphs_per_sec", "value": rate, "unit": "graphs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample_graphs_per_step": n},
            "cpu_baseline": {"value": rate, "unit": "graphs/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def run_b200(args):
    import torch
    import torch.distributed as dist
    import gcn_string_b200 as g
    from gcn_string_b200 import _lib, synthetic
    from gcn_string_b200.distributed import DataParallelTrainer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout for the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()

    B, K, W = args.batch_graphs, args.steps, args.warmup
    ds = synthetic.make_dataset(args.pool_batches * B, seed=100000 * rank, n_mean=N_MEAN, deg=DEG, n_feat=N_FEAT)
    np.random.seed(1234 + rank)
    sched = g.optimizers.schedules.PiecewiseConstantDecay([0, 1], [0.02, 0.002, 0.0002])   # gcn.py:321-325

    def make(device_resident, shuffle, **store_kw):
        loader = g.DisjointLoader(ds, batch_size=B, epochs=None, shuffle=shuffle, symmetric=True,
                                  device_resident=device_resident, **store_kw)   # each rank owns its pool: shards are per-rank here
        model = g.GeneralGNN(CLASSES, activation="softmax", hidden=HIDDEN, message_passing=LAYERS, seed=0)
        model.build(N_FEAT)
        trainer = DataParallelTrainer(model, g.optimizers.SGD(learning_rate=sched), sync_bn=args.sync_bn)
        return loader, model, trainer

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(loop_body, n_steps):
        """CUDA-event time of n_steps, max over ranks (ms)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_steps):
            loop_body()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ------------------------------------------------ device-resident arm (`value`)
    loader, model, trainer = make(True, True)
    stats = {}

    def step():
        (x, a, i), y = next(loader)
        stats["n"], stats["nnz"] = a.n_rows, a.nnz
        trainer.train_step((x, a, i), y)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                   # before the warm-up: its start-up must not sit in the timed region
        sampler.wait_ready()
    for _ in range(W):
        step()
    launches0 = lib.gcs_debug_launch_count()
    sampler.mark_begin()
    ms_total = timed(step, K)
    sampler.mark_end()
    launches = lib.gcs_debug_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / K
    value = world * B * K / (ms_total / 1e3)

    # forward-only rate (BASELINE cfg2: inference forward, batch resident in HBM)
    (xf, af, i_f), _yf = next(loader)

    def fwd():
        model((xf, af, i_f), training=False)
    for _ in range(3):
        fwd()
    n_fwd = max(5, K)
    ms_fwd = timed(fwd, n_fwd) / n_fwd

    def fwd_batched():                       # the same with shuffling + device batching of every batch inside the timing
        (x, a, i), y = next(loader)
        model((x, a, i), training=False)
    for _ in range(2):
        fwd_batched()
    ms_fwd_batched = timed(fwd_batched, max(3, K // 2)) / max(3, K // 2)

    # per-op device times of two more steps (CUDA events around every op, same stream)
    _lib.profile_begin()
    step()
    step()
    torch.cuda.synchronize()
    prof = _lib.profile_end()
    n_nodes, nnz = stats["n"], stats["nnz"]
    hbm_peak, tensor_peak, peak_src = _peaks()
    roofline, roofline_tensor, ops_report = None, None, {}
    tot = sum(ms for _, ms in prof.values()) or 1.0
    for label, (cnt, ms) in prof.items():
        ops_report[label] = {"launches_timed": cnt, "ms_per_call": ms / cnt, "share_of_step": ms / tot}
    if "spmm_fwd" in prof:
        cnt, ms = prof["spmm_fwd"]
        t = ms / cnt / 1e3
        alg = 4.0 * n_nodes * HIDDEN * 2 + 4.0 * nnz + 4.0 * (n_nodes + 1)
        ach = alg / t / 1e9
        roofline = {"kernel": "spmm_rb4_kernel<true> (K3, GeneralConv aggregation fwd, BN+PReLU fused on load)",
                    "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "peak_source": peak_src + " (MEASURED_PEAKS.json hbm_gbs)", "traffic": _ncu_traffic(),
                    "traffic_source": "profiles/r01_spmm_rb4_ncu.json (dram__bytes_read.sum + dram__bytes_write.sum, one launch)",
                    "algorithmic_bytes_per_launch": alg, "us_per_launch": t * 1e6,
                    "edges_per_sec": nnz / t, "frac_of_nominal_8TBs": ach / 8000.0}
    # the time-dominant kernels are the dense transforms (error-compensated fp16 split on tcgen05): fp32-equivalent rate of the forward GEMMs
    if "linear_fwd" in prof:
        cnt, ms = prof["linear_fwd"]
        steps_prof = 2
        flops = 2.0 * n_nodes * (N_FEAT * HIDDEN + HIDDEN * HIDDEN + HIDDEN * HIDDEN * sum(range(1, LAYERS + 1)))
        t = ms / steps_prof / 1e3
        ach = flops / t / 1e12
        roofline_tensor = {"kernel": "linear_tc_pair_kernel<f16> (K1, forward dense transforms; 3 fp16 tcgen05 MMAs per fp32 "
                                     "product: hi*hi + hi*lo + lo*hi of fp16-split operands)",
                           "bound": "tensor", "achieved": ach, "achieved_executed_f16": 3.0 * ach, "peak": tensor_peak,
                           "unit": "TFLOP/s", "frac": 3.0 * ach / tensor_peak,
                           "peak_source": peak_src + " (MEASURED_PEAKS.json bf16_tflops_sustained: kind::f16 runs at the bf16 "
                                          "rate; sustained figure, the GEMMs run back to back under the power cap)",
                           "flops_per_step_fwd": flops, "ms_per_step_fwd_gemms": t * 1e3,
                           "note": "frac counts the three fp16 passes as executed work; the kernel is bound by the L2->SM "
                                   "ingest of the raw fp32 activations (32 of 48 KB per 64-wide K block), not by the tensor "
                                   "pipe; the first layer (K=32, tf32 split) and the pooled post-MLP are included in the time"}
    del loader, model, trainer
    torch.cuda.empty_cache()

    # ------------------------------------------------ host-resident arm (`e2e`)
    loader_h, model_h, trainer_h = (make(False, True, device_gather=True) if args.e2e_shuffle else make(False, False))
    io = {"k": 0}
    host_loss = [torch.empty(2, dtype=torch.float32).pin_memory() for _ in range(2)]
    copied = [torch.cuda.Event(), torch.cuda.Event()]

    def step_e2e():
        # every step: the batch's graphs are copied out of pinned host memory (H2D, cudaMemcpyAsync of the slices), the
        # batching kernel and the train step run, {loss, acc} go back to a pinned host buffer (D2H); the host READS the
        # result of step t-1 while step t runs, so that the read does not drain the GPU queue (the copies of all K
        # steps are inside the timed region)
        k = io["k"]
        (x, a, i), y = next(loader_h)
        loss_acc, _ = trainer_h.train_step((x, a, i), y)
        host_loss[k % 2].copy_(loss_acc, non_blocking=True)
        copied[k % 2].record()
        if k > 0:
            copied[(k - 1) % 2].synchronize()
            io["loss"] = float(host_loss[(k - 1) % 2][0])
        io["h2d"] = loader_h.store.h2d_bytes_last
        io["k"] = k + 1
    for _ in range(W):
        step_e2e()
    ms_e2e = timed(step_e2e, K)
    e2e_value = world * B * K / (ms_e2e / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n = 32
        rate, threads, sec, _ = cpu_reference_rate(n, 3, 1)
        cpu = {"value": rate, "unit": "graphs/s", "cores": threads, "kind": "port",
               "sample": f"{n} graphs/step (BASELINE cfg1) of the same workload through the CPU restatement of the "
                         f"reference op sequence (scipy collate + PyTorch-CPU fp32 fwd/bwd/SGD), median of 3 steps, "
                         f"{sec * 1e3:.0f} ms/step"}
    line = {
        "metric": "train_graphs_per_sec", "value": value, "unit": "graphs/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "graphs_per_gpu_per_step": B, "nodes_per_step_per_gpu": n_nodes,
                   "nnz_per_step_per_gpu": nnz, "optimizer": "SGD PiecewiseConstantDecay (gcn.py:321-325)",
                   "parallelism": f"graph-sharded data parallel x{world}, one flat NCCL all-reduce (4.26 MB)/step",
                   "l2": "activations per step (cat 2.6 GB, h 0.5 GB/layer) far exceed the 126 MB L2; batches reshuffled "
                         "every step", "bn": ("synchronised BatchNorm statistics (16 extra fp64 all-reduces of <= 3H+1 values per step)"
                          if args.sync_bn and world > 1 else "replica-local BatchNorm statistics")},
        "e2e": {"value": e2e_value, "unit": "graphs/s", "ms_per_step": ms_e2e / K, "h2d_bytes_per_step": int(io["h2d"]),
                "d2h_bytes_per_step": 8, "note": ("dataset in pinned host memory, reshuffled batches; per step: a device kernel gathers "
                         "the batch's packed graphs out of the pinned arrays (H2D), device batching, train step, D2H of {loss, acc} "
                         "into a pinned buffer that the host reads one step later") if args.e2e_shuffle else
                "dataset in pinned host memory, consecutive batches; per step: H2D of the batch's "
                "packed graphs (cudaMemcpyAsync from the pinned arrays), device batching, train step, D2H of {loss, acc} into a "
                "pinned buffer that the host reads one step later"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_tensor": roofline_tensor,
        "cpu_baseline": cpu,
        "fwd_graphs_per_sec": world * B / (ms_fwd / 1e3), "fwd_ms_per_step": ms_fwd,
        "fwd_with_batching_ms_per_step": ms_fwd_batched,
        "edges_per_sec_train_step": world * nnz / (ms_step / 1e3),
        "ops": ops_report,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def _claim_stdout():
    """Keep stdout for the ONE JSON line: everything else any library writes to file descriptor 1 (NCCL's version
    banner lands there on some launches) is sent to stderr for the duration of the run."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch-graphs", type=int, default=B_GRAPHS)
    ap.add_argument("--pool-batches", type=int, default=4, help="synthetic pool size in batches per rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-shuffle", action="store_true", help="e2e arm with reshuffled batches: the selected graphs are "
                    "gathered out of pinned host memory by a device kernel (gcs_gather_graphs) instead of sliced copies")
    ap.add_argument("--sync-bn", action="store_true", help="all-reduce the BatchNorm statistics too (a G-GPU step then "
                    "equals one step on the union batch); off by default: the gradient all-reduce is the only collective")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    return run_reference(args) if args.impl == "reference" else run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
