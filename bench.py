#!/usr/bin/env python
"""bench.py -- cluster-tracking hot path throughput on B200 (driver contract, see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames F]

A "step" is one pass of the hot path over one synthetic Waymo-shaped sequence (BASELINE.json configs[1]):
0.08 m pick-one subsample voxelization -> sequence-level ground removal -> 3-radius neighbour graphs +
connected-component cluster proposals, through the reference-facing plugin (SimpleReg.forward with the
GroundPlaneRemover and ClusterProposal preprocessors of cluster_tracking_TLS_multiradius_every8.yaml).

  value  frames/s with the sequence already resident in HBM (CUDA events, max over ranks)
  e2e    frames/s through the same plugin call starting from pinned HOST buffers, host->device copies and the
         device->host read of the result inside the timed region
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md

N > 1 (torchrun, one rank per GPU): every rank processes its own sequence (replicas, "weak" scaling) -- the
path has no data-path collective at sequence granularity.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "cluster-tracking frames/sec (Waymo-shape seq)"
UNIT = "frames/s"
WORKLOAD = "ground removal + 0.08m voxelization + multi-radius graph + CC proposals, 198-frame synthetic sequence"
REFERENCE_BUDGET_S = 150.0  # wall-clock cap of the timed loop of --impl reference
POINT_KEYS = ["point_bxyz", "point_sweep", "point_feat", "segmentation_label", "instance_label"]


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("radius_search_bytes_per_launch")
        except Exception:
            return None
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """SM clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe).

    Default: an in-process NVML thread (two light queries every 50 ms).  A looping `nvidia-smi` child is the fallback
    (PCS_BENCH_CLOCKS=smi): its full per-iteration query holds driver locks long enough to show up as occasional
    +30..150 ms steps on this host-synchronising path."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.thread = gpu_index, None, None
        self.mode = os.environ.get("PCS_BENCH_CLOCKS", "nvml")
        self.sm, self.mx, self.reasons, self.stop = [], 0.0, set(), False

    def _nvml_handle(self):
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def _nvml_loop(self, nv, h):
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown")
                else nv.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown",
                                               getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown",
                                               getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap",
                                        getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4))}
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons",
                             getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None))
        while not self.stop:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                if reasons_fn is not None:
                    r = int(reasons_fn(h))
                    for name, bit in bits.items():
                        if r & int(bit):
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def start(self):
        if self.mode == "off":
            return
        if self.mode == "nvml":
            try:
                nv, h = self._nvml_handle()
                self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
                import threading
                self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
                self.thread.start()
                return
            except Exception:
                self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def summary(self):
        if self.thread is not None:
            self.stop = True
            self.thread.join(timeout=2)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx or None,
                    "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml"}
        lines = []
        if self.proc is not None:
            try:
                self.proc.terminate()
                out, _ = self.proc.communicate(timeout=5)
                lines = [ln for ln in out.splitlines() if ln.strip()]
            except Exception:
                pass
        sm, mx, reasons = [], 0.0, set()
        for ln in lines:
            s = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(s[1]))
                mx = max(mx, float(s[2]))
            except Exception:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


def build_model(device):
    from pcseqlearning_b200.config import cluster_tracking_cfg
    from pcseqlearning_b200.simple_reg import SimpleReg
    cfg = cluster_tracking_cfg(out_dir="/tmp/pcseq_bench_out")
    cfg.PREPROCESSORS = [p for p in cfg.PREPROCESSORS if p.NAME in ("GroundPlaneRemover", "ClusterProposal")]
    for p in cfg.PREPROCESSORS:
        p.VERBOSE = False
        p.USE_CACHE = False  # never reuse pillar_height.pth: every step recomputes the ground field
        p.LOG_DIR = None
        p.EVALUATE = False  # GT-IoU bookkeeping is evaluation, not the algorithm
    cfg.SAVE_DIR = None
    model = SimpleReg(cfg, {}, None).to(device)
    model.train()
    return model


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from pcseqlearning_b200 import ops
    from pcseqlearning_b200.data_staging import DevicePrefetcher
    from pcseqlearning_b200.synthetic import generate_sequence

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # connect the communicator now: NCCL's lazy set-up (proxy threads, channel buffers) otherwise lands on the
        # first timed steps of every rank
        t = torch.zeros(1, device=dev)
        dist.all_reduce(t)
        dist.barrier()
        torch.cuda.synchronize()
    t0 = time.time()
    # every rank processes the SAME synthetic sequence (identical work per GPU: replicas / weak scaling)
    cache = args.cache and f"{args.cache}.f{args.frames}.pt"
    if cache and os.path.exists(cache):
        batch = torch.load(cache, map_location=dev, weights_only=False)
    else:
        batch = generate_sequence(0, num_frames=args.frames, device=dev)
        if cache and rank == 0:
            torch.save(batch, cache)
    torch.cuda.synchronize()
    n_points = int(batch["point_bxyz"].shape[0])
    gen_s = time.time() - t0
    model = build_model(dev)
    # pinned host copies of the per-point inputs for the end-to-end leg
    host = {k: batch[k].cpu().pin_memory() for k in POINT_KEYS}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    def step_device():
        model(batch)
        return model.forward_dict["sequences"][0]

    def result_of(seq):
        return torch.stack([seq["num_component_rad1x25"].sum(), seq["num_component_rad0x75"].sum(),
                            seq["num_component_rad0x25"].sum()]).cpu()  # device -> host read of the result

    def host_batches(n):
        for _ in range(n):
            b = dict(batch)
            b.update(host)  # the per-point inputs start in pinned host memory every step
            yield b

    def run_e2e(steps):
        """The user-facing loop: DevicePrefetcher (copy of batch k+1 overlaps the processing of batch k) ->
        SimpleReg.forward -> result read back on the host.  All `steps` H2D copies happen inside the loop."""
        seq = out = None
        for b in DevicePrefetcher(host_batches(steps), dev):
            seq = out = None
            model(b)
            seq = model.forward_dict["sequences"][0]
            out = result_of(seq)
        return seq, out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_ms = []

    def timed(fn, steps):
        # like timeit: no cyclic-GC pass inside the timed loop (a full collection of the torch-sized heap is a
        # 100-200 ms host stall that lands on one early step and never again)
        gc.collect()
        gc.disable()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        a.record()
        for i in range(steps):
            torch.cuda.nvtx.range_push("pcs_step")  # lets ncu restrict a launch list to the timed region
            t0 = time.perf_counter()
            res = None  # the previous step's result is released first (the trainer does not keep it either); holding
            res = fn()  # it makes the caching allocator cudaMalloc ~1 GB mid-run, a 100-200 ms stall on one step
            torch.cuda.nvtx.range_pop()
            marks[i].record()
            if os.environ.get("PCS_BENCH_DIAG"):
                st = torch.cuda.memory_stats()
                print("[diag] step %d host %.1f ms, cudaMalloc calls %d, reserved %.2f GB, retries %d" % (
                    i, (time.perf_counter() - t0) * 1e3, st.get("num_device_alloc", -1),
                    st.get("reserved_bytes.all.current", 0) / 1e9, st.get("num_alloc_retries", -1)),
                    file=sys.stderr, flush=True)
        b.record()
        barrier()
        gc.enable()
        ms = a.elapsed_time(b)
        step_ms.clear()
        prev = a
        for m in marks:
            step_ms.append(round(prev.elapsed_time(m), 2))
            prev = m
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, res

    for _ in range(max(args.warmup, 3)):
        seq = None
        seq = step_device()
    seq = None
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ops.enable_event_log(True)
    ops.reset_launch_count()
    ms_dev, seq = timed(step_device, args.steps)
    dev_step_ms = list(step_ms)
    launches = ops.launch_count()
    log = ops.event_log()
    ops.enable_event_log(False)
    run_e2e(4)  # warm-up of the staged loop (same allocation pattern as the timed one)
    ms_e2e, (seq_e, out_e) = timed(lambda: run_e2e(args.steps), 1)
    clocks = sampler.summary()

    # roofline of the dominant kernel (radius search): algorithmic bytes per launch / mean launch duration
    peak, peak_kind = peaks()
    ev = log.get("radius_search", [])
    durs = [a.elapsed_time(b) for a, b, _ in ev]
    n_q = ev[0][2]["n_query"] if ev else 0
    # B_rg = 16*N_ref + 16*N_q + 4*N_q + e*E with e = 0: the fused kernel consumes the lists in-kernel
    alg_bytes = (16 + 16 + 4) * n_q
    mean_ms = sum(durs) / max(len(durs), 1)
    achieved = alg_bytes / (mean_ms * 1e-3) / 1e9 if durs else 0.0
    hb = log.get("hash_build", [])
    hb_ms = sum(a.elapsed_time(b) for a, b, _ in hb) / max(len(hb), 1)
    hb_n = hb[0][2]["n"] if hb else 0

    frames_total = args.frames * world
    value = frames_total / (ms_dev / args.steps / 1e3)
    e2e = frames_total / (ms_e2e / args.steps / 1e3)
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_dev / args.steps, 3), "step_ms": dev_step_ms,
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames": args.frames, "points_per_sequence": n_points,
                   "points_after_subsample": int(seq["full_point_fxyz"].shape[0]),
                   "points_after_ground_removal": int(seq["point_fxyz"].shape[0]),
                   "points_per_s": round(n_points * world / (ms_dev / args.steps / 1e3)),
                   "l2": "inputs (%.0f MB per step) larger than L2" % (h2d_bytes / 1e6),
                   "parallelism": "replicas" if world > 1 else "single", "generate_s": round(gen_s, 1)},
        "e2e": {"value": round(e2e, 3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": int(out_e.numel() * out_e.element_size()),
                "ms_per_step": round(ms_e2e / args.steps, 3),
                "pipeline": "pinned host batch -> DevicePrefetcher (copy of step k+1 overlaps step k) -> "
                            "SimpleReg.forward -> .cpu() of the component counts"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "radius_search_kernel<fused union-find>", "achieved": round(achieved, 2),
                     "peak": peak, "peak_kind": peak_kind, "unit": "GB/s", "frac": round(achieved / peak, 5),
                     "traffic": measured_traffic(), "launches_timed": len(durs), "mean_launch_ms": round(mean_ms, 4),
                     "launch_ms": [round(d, 3) for d in durs[:6]],
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "hash_build": {"achieved": round(hb_n * 28 / (hb_ms * 1e-3) / 1e9, 2) if hb else None,
                                    "mean_ms": round(hb_ms, 4), "algorithmic_bytes_per_launch": hb_n * 28}},
        "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(batch, args, sample_frames=args.cpu_frames)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def oracle_pipeline(point_fxyz_host, sample_frames):
    """The same hot path restated on the CPU (oracle/), on the first `sample_frames` frames."""
    import numpy as np
    from oracle import cpu_ops, ground_np
    pts = point_fxyz_host[point_fxyz_host[:, 0] < sample_frames]
    t = {}
    t0 = time.perf_counter()
    pick = cpu_ops.subsample_pick(pts)
    pts = np.ascontiguousarray(pts[pick])
    t["subsample"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    cfg = dict(PILLAR_SIZE=[2, 2], LR=0.01, DECAY_STEPS=[1600], RIGID_WEIGHT=0.5, MAX_NUM_ITERS=10000,
               TRUNCATE_HEIGHT=[0.5], RANSAC=True, SIGMA2=0.0025, JointOpt=True, K=8)
    height = ground_np.ground_plane_removal(pts, cfg)[0]
    pts = np.ascontiguousarray(pts[~(height < 0.5)])
    t["ground"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    ncomp = []
    for r in (1.25, 0.75, 0.25):
        _, n = cpu_ops.propose_clusters(pts, r)
        ncomp.append(n)
    t["graphs_cc"] = time.perf_counter() - t0
    return t, ncomp


def cpu_baseline(batch, args, sample_frames):
    import torch
    from oracle import cpu_ops
    fx = torch.cat([batch["point_sweep"].reshape(-1, 1).float(), batch["point_bxyz"][:, 1:]], -1).cpu().numpy()
    cores = len(os.sched_getaffinity(0))
    cpu_ops.set_threads(cores)
    torch.set_num_threads(cores)
    t0 = time.perf_counter()
    stages, _ = oracle_pipeline(fx, sample_frames)
    dt = time.perf_counter() - t0
    return {"value": round(sample_frames / dt, 4), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {sample_frames} of {args.frames} frames of the same sequence through oracle/ "
                      f"(subsample + ground + 3-radius graph + CC; the per-sequence ground solve is amortised over "
                      f"{sample_frames} frames only)",
            "stage_s": {k: round(v, 2) for k, v in stages.items()}}


def run_reference(args, rank, world):
    """--impl reference: the reference has no CPU implementation of this path and its Python cannot be installed
    here (torch_scatter / torch_cluster / torch_geometric are absent); the arm times the CPU restatement in
    oracle/ (pinned against the reference's own code, see oracle/README.md) on all host cores."""
    if rank != 0:
        return
    import numpy as np
    import torch
    from oracle import cpu_ops
    from pcseqlearning_b200.synthetic import generate_sequence
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    sample = args.cpu_frames
    batch = generate_sequence(0, num_frames=sample, device=dev)
    fx = torch.cat([batch["point_sweep"].reshape(-1, 1).float(), batch["point_bxyz"][:, 1:]], -1).cpu().numpy()
    n_points = fx.shape[0]
    cores = len(os.sched_getaffinity(0))
    cpu_ops.set_threads(cores)
    torch.set_num_threads(cores)
    for _ in range(min(args.warmup, 1)):
        oracle_pipeline(fx, sample)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        stages, _ = oracle_pipeline(fx, sample)
        done += 1
        if time.perf_counter() - t0 > REFERENCE_BUDGET_S:  # keep the whole arm within a few minutes
            break
    dt = (time.perf_counter() - t0) / done
    value = sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world,
        "steps": done, "warmup": min(args.warmup, 1), "ms_per_step": round(dt * 1e3, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames": args.frames, "points_per_step": n_points,
                   "note": f"each step = a bounded sample ({sample} frames) of the workload on the host CPU; "
                           f"{done} of the {args.steps} requested steps fit the {REFERENCE_BUDGET_S:.0f} s budget"},
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} frames per step through oracle/ (CPU restatement of the reference)",
                         "stage_s": {k: round(v, 2) for k, v in stages.items()}},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=198)
    ap.add_argument("--cpu-frames", type=int, default=2, dest="cpu_frames")
    ap.add_argument("--cache", default=None, help="path prefix to cache the synthetic sequence (profiling runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true", dest="no_cpu")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
