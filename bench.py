#!/usr/bin/env python
"""bench.py — the headline benchmark of the D-VQVAE hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

Workload (BASELINE.json configs[1]): fused VQ lookup, N = 4 194 304 fp32 latents PER GPU,
e_dim = 64, n_e = 512, train path (indices + z_q + loss + perplexity + usage histogram),
min_encodings lazy (not materialised).  A "step" is one pass of the path over that batch.
Weak scaling: every rank holds its own 4M-row shard, codebook replicated, one all-reduce of
(hist, sse) per step.  Prints ONE JSON line (rank 0).

  value    whole-job latents/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e      same metric through the host-buffer C-ABI call (dvq_vq_forward_host): pinned host
           z -> H2D -> kernels -> D2H z_q + idx, copies inside the timed region
  roofline dominant kernel vs the measured HBM peak (MEASURED_PEAKS.json), algorithmic bytes
           N*(8*D+8) + 4*K*D per launch, duration from CUDA events recorded
           around that kernel inside the library on the launching stream
  cpu_baseline  the reference's torch CPU op sequence (oracle/ref_port_torch.py, kind "port")
           on a bounded sample of the same workload, all host threads
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
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "d-vqvae_b200"))

N_PER_GPU = 4194304
E_DIM = 64
N_E = 512
E2E_CHUNK_ROWS = 262144          # rows per chunk of the host-buffer pipeline (e2e leg)
AL, BETA = 1.0, 0.25
METRIC = "vq_latents_per_sec"
UNIT = "latents/s"
WORKLOAD = ("BASELINE config 2: fused VQ lookup, N=4194304 fp32 latents per GPU, e_dim=64, n_e=512, "
            "istrain=True (idx + z_q + loss + perplexity + histogram), min_encodings lazy")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 0.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 100 ms while the benchmark runs; the median SM
    clock is taken over the samples drawn under load (power above idle)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        pw = []
        for r in self.rows:
            try:
                pw.append(float(r[3]))
            except Exception:
                pass
        load_floor = (min(pw) + 0.25 * (max(pw) - min(pw))) if pw else 0.0
        for r in self.rows:
            try:
                if float(r[3]) < load_floor:
                    continue
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_rate(target_seconds: float, max_rows: int):
    """Reference op sequence on host cores (oracle/ref_port_torch.py), bounded sample."""
    import torch
    from oracle import ref_port_torch as port
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(2000)
    E = (torch.rand(N_E, E_DIM, generator=g) * 2 - 1) / N_E
    chunk = 65536
    z = torch.randn(chunk, E_DIM, generator=g)
    port.quantize_rows(z, E, AL, BETA, True)                      # warm-up
    t0 = time.perf_counter()
    port.quantize_rows(z, E, AL, BETA, True)
    per_chunk = max(time.perf_counter() - t0, 1e-4)
    n_chunks = int(max(1, min(max_rows // chunk, target_seconds / per_chunk)))
    zs = torch.randn(n_chunks * chunk, E_DIM, generator=g)
    t0 = time.perf_counter()
    port.quantize_chunked(zs, E, AL, BETA, True, chunk)
    dt = time.perf_counter() - t0
    rows = n_chunks * chunk
    return rows / dt, torch.get_num_threads(), "%d rows (%d chunks of 65536) of the config-2 workload, train path, %.1f s" % (rows, n_chunks, dt)


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (torch CPU port of
    network/vqvae/quantizer.py — the Python reference cannot travel to the GPU box)."""
    if rank != 0:
        return
    import torch
    from oracle import ref_port_torch as port
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(2000)
    E = (torch.rand(N_E, E_DIM, generator=g) * 2 - 1) / N_E
    chunk = 65536
    z1 = torch.randn(chunk, E_DIM, generator=g)
    port.quantize_rows(z1, E, AL, BETA, True)
    t0 = time.perf_counter()
    port.quantize_rows(z1, E, AL, BETA, True)
    per_chunk = max(time.perf_counter() - t0, 1e-4)
    budget = 150.0 / max(1, args.steps + args.warmup)            # whole run within a few minutes
    n_chunks = int(max(1, min(N_PER_GPU // chunk, budget / per_chunk)))
    rows = n_chunks * chunk
    z = torch.randn(rows, E_DIM, generator=g)
    for _ in range(args.warmup):
        port.quantize_chunked(z, E, AL, BETA, True, chunk)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        port.quantize_chunked(z, E, AL, BETA, True, chunk)
    dt = time.perf_counter() - t0
    value = rows * args.steps / dt
    sample = "%d of %d rows per step (%d chunks of 65536), train path" % (rows, N_PER_GPU, n_chunks)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample, "device": "host CPU"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer leg (default: min(steps, 10))")
    ap.add_argument("--path", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-chunk-rows", type=int, default=E2E_CHUNK_ROWS, help="rows per chunk of the host-buffer pipeline")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as tdist
    import dvq
    from dvq import _cabi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        tdist.init_process_group("nccl", device_id=dev)

    gen = torch.Generator(device=dev).manual_seed(2000 + rank)
    cb_gen = torch.Generator(device=dev).manual_seed(2000)       # replicated codebook
    vq = dvq.VectorQuantizer(N_E, E_DIM, BETA, AL).to(dev)
    with torch.no_grad():
        vq.embedding.weight.copy_((torch.rand(N_E, E_DIM, device=dev, generator=cb_gen) * 2 - 1) / N_E)
    vq.path = {"auto": _cabi.DVQ_PATH_AUTO, "simt": _cabi.DVQ_PATH_SIMT, "tc": _cabi.DVQ_PATH_TC}[args.path]
    vq.onehot_limit_bytes = 0
    if world > 1:
        dvq.dist.shard_module(vq)
    z = torch.randn(N_PER_GPU, E_DIM, device=dev, generator=gen)

    def step():
        with torch.no_grad():
            return vq(z, True)

    def fence():
        torch.cuda.synchronize(dev)
        if world > 1:
            tdist.barrier()
            torch.cuda.synchronize(dev)

    # clocks / throttle reasons are sampled from the first warm-up step to the end of the e2e leg (the
    # device-resident timed region alone can be shorter than one 100 ms sampling period)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        out = step()
    fence()

    # ---- timed region: device-resident inputs --------------------------------------------------
    _cabi.lib.dvq_profile_enable(1)
    launches0 = _cabi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    ev0.record()
    for _ in range(args.steps):
        out = step()
    ev1.record()
    fence()
    # the library's per-stage CUDA events (recorded on the launching stream), mean over the timed steps
    prof_ms, prof_cnt = _cabi.profile_mean()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = _cabi.launch_count() - launches0
    _cabi.lib.dvq_profile_enable(0)
    if world > 1:
        t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    loss, z_q, ppl, enc, idx = out
    value = N_PER_GPU * world * args.steps / (elapsed_ms * 1e-3)

    # ---- e2e: host buffers through dvq_vq_forward_host --------------------------------------------
    e2e_steps = args.e2e_steps or min(args.steps, 10)
    hq = dvq.HostQuantizer(chunk_rows=args.e2e_chunk_rows, n_e_max=N_E, e_dim_max=E_DIM, device=dev)
    z_host = torch.empty((N_PER_GPU, E_DIM), dtype=torch.float32, pin_memory=True)
    z_host.copy_(z)
    E_host = vq.embedding.weight.detach().cpu().pin_memory()
    zq_host = torch.empty((N_PER_GPU, E_DIM), dtype=torch.float32, pin_memory=True)
    idx_host = torch.empty((N_PER_GPU, 1), dtype=torch.int64, pin_memory=True)
    for _ in range(2):
        hq.forward(z_host, E_host, True, AL, BETA, out_zq=zq_host, out_idx=idx_host, path=vq.path)
    fence()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        l_h, _, p_h, _ = hq.forward(z_host, E_host, True, AL, BETA, out_zq=zq_host, out_idx=idx_host, path=vq.path)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = N_PER_GPU * world * e2e_steps / e2e_s
    same_idx = bool(torch.equal(idx_host, idx.cpu()))
    hq.close()
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        hbm_gbs, bf16_tf, peak_src = measured_peaks()
        alg_bytes = N_PER_GPU * (8 * E_DIM + 8) + 4 * N_E * E_DIM
        k_ms = prof_ms[1]
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s", "frac": achieved / hbm_gbs,
                "traffic": None, "kernel": "vq main kernel (stage 1 of dvq_vq_forward)", "kernel_ms": k_ms,
                "refine_ms": prof_ms[2], "launches_averaged": prof_cnt[1], "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes,
                "tensor_frac_of_half_bf16_peak": (2.0 * N_PER_GPU * N_E * E_DIM / (k_ms * 1e-3) / 1e12) / (bf16_tf / 2) if bf16_tf else None}
        tr = os.path.join(ROOT, "profiles", "traffic_r01.json")
        if os.path.exists(tr):
            try:
                roof["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch")
            except Exception:
                pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, cores, sample = cpu_port_rate(15.0, N_PER_GPU)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "l2_policy": "inputs (1 GiB z per GPU) larger than the 126 MB L2; no flush needed",
                       "path": args.path, "parallelism": "rows sharded over %d GPU(s), codebook replicated, 1 all-reduce of (hist,sse)" % world,
                       "loss": float(loss.item()), "perplexity": float(ppl.item())},
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": N_PER_GPU * E_DIM * 4 + N_E * E_DIM * 4,
                    "d2h_bytes_per_step": N_PER_GPU * E_DIM * 4 + N_PER_GPU * 8 + 8, "steps": e2e_steps,
                    "ms_per_step": e2e_s / e2e_steps * 1e3, "api": "dvq_vq_forward_host (pinned host buffers, 3-stream chunk pipeline, %d-row chunks)" % args.e2e_chunk_rows,
                    "indices_equal_device_path": same_idx},
            "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
