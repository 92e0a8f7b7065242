#!/usr/bin/env python
"""bench.py -- tracks/sec of the BoT-SORT per-frame step (Kalman + IoU + ReID distance + LAP +
lifecycle) on synthetic streams, B200 vs the CPU restatement of the reference.

    python bench.py --gpus 1 --steps 20 --warmup 3            # our arm (N=1)
    torchrun ... bench.py --gpus N --steps K --warmup W       # one rank per GPU, one stream per rank
    python bench.py --impl reference --steps K --warmup W     # CPU arm (oracle port of the reference)

One "step" = one bt_update_arrays call = one BoTSORT.update (demo:1291-1639) on one frame of
`n` synthetic detections with 2048-d features against `n` live tracks (steady state: every
track is matched).  Prints ONE JSON line (rank 0).
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[2]: the configuration the north-star target is quoted on
    "c3": dict(n=2000, feat_dim=2048, reid=True, desc="2000 tracks x 2000 dets, fused IoU+cosine(2048-d), 1 stream/GPU"),
    # BASELINE.json configs[1]
    "c2": dict(n=512, feat_dim=2048, reid=False, desc="512 tracks x 512 dets, IoU-only association, 1 stream/GPU"),
    # BASELINE.json configs[0] (CPU-runnable case)
    "c1": dict(n=64, feat_dim=2048, reid=True, desc="64 tracks x 64 dets, 2048-d ReID features, 1 stream/GPU"),
    # per-stream shape of BASELINE.json configs[4]
    # BASELINE.json configs[3]: detector side chained on the device into the tracker
    "c4": dict(n=40, feat_dim=2048, reid=True, persons=40, img_h=480, img_w=640,
               desc="480x640 synthetic frames + raw YOLOX head (6300 anchors x 4 classes) -> decode+NMS+crop gather -> "
                    "stub ReID encoder (256x128 crops -> 2048-d fp16) -> BoTSORT.update, all chained on the device, 1 stream/GPU"),
    "c5": dict(n=1000, feat_dim=2048, reid=True, streams=4,
               desc="1000 tracks x 1000 dets per stream, 4 independent streams per GPU (BASELINE config 5: 32 streams over 8 GPUs)"),
}
METRIC = "tracks/sec (Kalman+IoU+ReID-dist+LAP per frame)"


def make_frames(wl, count, seed):
    from botsort_b200.synthetic import SceneConfig, SyntheticScene
    scene = SyntheticScene(SceneConfig(n_ids=wl["n"], feat_dim=wl["feat_dim"], seed=seed, with_features=wl["reid"]))
    return [scene.next_frame() for _ in range(count)]


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2]))
            except ValueError:
                continue
            for name, val in zip(names, p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference; also the cpu_baseline leg of our arm)
# ------------------------------------------------------------------------------------------------
def config_of(wl, streams_per_gpu=1):
    """The `config` object of the JSON line -- the SAME keys and values in both arms (GPU and reference)."""
    return {"workload": wl["desc"], "tracks": wl["n"], "dets": wl["n"], "feat_dim": wl["feat_dim"],
            "streams_per_gpu": streams_per_gpu,
            "l2": "GPU arm: L2 flushed between timed steps (256 MiB memset, untimed); per-step CUDA events on the ctx stream, summed",
            "sharding": "independent video streams block-partitioned over ranks, no data-path collective"}


def cpu_frame_times(wl, frames, mode):
    """Run the oracle tracker over `frames` (first one = birth frame, untimed) and return per-frame wall
    times of the FULL frame step -- nothing is sampled or extrapolated.  mode 'faithful' keeps the
    reference's structure (pure-Python IoU double loop over every (track, detection) pair, demo:1742; one
    Kalman update per match); mode 'vectorized' is the broadcast-IoU / batched-update restatement."""
    from oracle import oracle_np as O
    trk = O.OracleBoTSORT(mode=mode, lap_solver="jv", use_features=wl["reid"])
    times = []
    for k, fr in enumerate(frames):
        feats = fr["feats"] if wl["reid"] else None
        t0 = time.perf_counter()
        trk.update_arrays(fr["boxes"], fr["scores"], feats)
        dt = time.perf_counter() - t0
        if k > 0:
            times.append(dt)
    return times


def _all_host_threads():
    try:    # torchrun exports OMP_NUM_THREADS=1: give NumPy/BLAS every host core back for the CPU legs
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port, the reference's own loop structure) on
    the host cores of this box, every frame in full.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    n = wl["n"]
    _all_host_threads()
    K, W = args.steps, args.warmup
    frames = make_frames(wl, 1 + W + K, seed=1234)
    t_wall0 = time.perf_counter()
    t_all = cpu_frame_times(wl, frames, "faithful")
    t = t_all[W:]
    ms = 1e3 * float(np.mean(t))
    value = n / (ms / 1e3)
    # the stronger CPU baseline beside it (a few frames, same scene)
    t_vec = cpu_frame_times(wl, frames[: 1 + min(len(frames) - 1, 1 + 4)], "vectorized")
    t_vec = t_vec[1:] if len(t_vec) > 1 else t_vec
    sample = (f"{len(t)} timed frames (+{W} warm-up) of the full frame step in the reference's loop structure: "
              f"pure-Python IoU over all {n}x{n} pairs, per-match Kalman update; nothing extrapolated")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "tracks/s", "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(wl, int(wl.get("streams", 1))),
        "timed_region_s": float(np.sum(t)), "run_wall_s": time.perf_counter() - t_wall0,
        "cpu_baseline": {"value": value, "unit": "tracks/s", "cores": os.cpu_count(), "kind": "port", "sample": sample,
                         "min_ms": 1e3 * float(np.min(t)), "median_ms": 1e3 * float(np.median(t)),
                         "vectorized_numpy": {"value": n / float(np.mean(t_vec)), "unit": "tracks/s",
                                              "sample": f"{len(t_vec)} frames, broadcast IoU + batched Kalman update"},
                         "note": "Python reference cannot travel to the GPU box; oracle/oracle_np.py (pinned against it "
                                 "in the build container) stands in. NumPy/BLAS may use all cores; the Python loops are "
                                 "single-threaded like the reference. LAP = C port of lap 0.4.0's JV (oracle/lapjv_port.c)."},
        "e2e": {"value": value, "unit": "tracks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import botsort_b200 as bs
    from botsort_b200._lib import BT_DEVICE, BT_F16, BT_F32, BT_HOST

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
    host_cores = None
    if local_world > 1 and hasattr(os, "sched_setaffinity"):
        # one rank per GPU on one host: every rank gets its own slice of the host cores NEXT TO ITS GPU (the
        # enqueue / bookkeeping thread is on the frame's critical path, and the pinned staging buffers are
        # first-touched on the node the GPU's DMA reads from)
        from botsort_b200.sharding import gpu_local_cpus, split_cores_numa_local
        topo = [gpu_local_cpus(r) for r in range(local_world)]
        host_cores = split_cores_numa_local(topo, sorted(os.sched_getaffinity(0)), local)
        try:
            os.sched_setaffinity(0, host_cores)
        except OSError:
            host_cores = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.workload]
    n, D, reid = wl["n"], wl["feat_dim"], wl["reid"]
    K, W = args.steps, max(args.warmup, 3)
    f16 = args.feat_dtype == "f16"
    np_feat, th_feat, bt_feat = (np.float16, torch.float16, BT_F16) if f16 else (np.float32, torch.float32, BT_F32)

    from botsort_b200.sharding import aggregate_throughput, max_over_ranks, shard_streams
    S = int(wl.get("streams", 1))                       # independent video streams on THIS GPU
    my_streams = shard_streams(S * world, world, rank)  # global stream ids of this rank (no data-path collective)
    cap = (n + n // 8 + 255) // 128 * 128
    # ONE ctx per GPU: its S video streams are a leading batch dimension of every kernel
    ctx = bs.Context(max_tracks=cap, max_dets=cap, feat_dim=D, device=local, n_streams=S)
    sids = list(range(S))
    cfg = ctx.default_config()
    cfg.with_reid = 1 if reid else 0
    cu_stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    # every stream has its own synthetic scene (weak scaling: each GPU tracks its own video streams)
    n_frames = 1 + W + K
    data = []
    for sid in my_streams:
        frames = make_frames(wl, n_frames, seed=1234 + sid)
        bh = [torch.from_numpy(f["boxes"]).pin_memory() for f in frames]
        sh = [torch.from_numpy(f["scores"]).pin_memory() for f in frames]
        # the ReID encoder's output dtype: fp16 is what the reference's TensorRT FastReID engine computes in
        # (demo:738, demo:35-49); --feat-dtype f32 feeds float32 rows (onnxruntime's output type)
        fh = [torch.from_numpy(f["feats"].astype(np_feat)).pin_memory() for f in frames] if reid else [None] * n_frames
        data.append(dict(bh=bh, sh=sh, fh=fh, bd=[b.to(dev) for b in bh], sd=[x.to(dev) for x in sh],
                         fd=[f.to(dev) for f in fh] if reid else [None] * n_frames))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def ptrs(i, where):
        b = [d["b" + where][i] for d in data]
        s = [d["s" + where][i] for d in data]
        f = [d["f" + where][i] for d in data]
        return ([x.data_ptr() for x in b], [x.data_ptr() for x in s], [x.data_ptr() if reid else 0 for x in f],
                [int(x.shape[0]) for x in b])

    def place_inputs(i):
        """Zero-copy ingest (bt_input_buffers): frame i is written into the ctx's own association
        buffers -- what a detector / ReID engine bound to these addresses does -- OUTSIDE the timed region."""
        out = []
        for k, d in enumerate(data):
            pb, ps, pf = ctx.input_buffers(k)
            m = int(d["bd"][i].shape[0])
            for src, ptr in ((d["bd"][i], pb), (d["sd"][i], ps)) + (((d["fd"][i], pf),) if (reid and f16) else ()):
                bs_ = src.numel() * src.element_size()
                torch.cuda.current_stream().synchronize()
                _cudart().cudaMemcpy(ptr, src.data_ptr(), bs_, 3)          # device to device
            out.append((pb, ps, (pf if f16 else d["fd"][i].data_ptr()) if reid else 0, m))
        return ([o[0] for o in out], [o[1] for o in out], [o[2] for o in out], [o[3] for o in out])

    def read_back():
        nbytes = 0
        for k in range(S):
            res = ctx.get_tracks(0, stream=k)               # ids + boxes of the returned list, on the host
            nbytes += res["tlbr"].nbytes + res["ids"].nbytes
        return nbytes

    def run_value_pass(profile=False, l2_flush=True):
        """Device-resident inputs (already in the association buffers when the timed region starts): frame 0 =
        births (untimed), W warm-up frames, then K timed frames; per-step CUDA events on the ctx stream."""
        ctx.tracker_reset(cfg)
        for i in range(0, 1 + W):
            b, s_, f, m = place_inputs(i)
            ctx.update_streams_raw(sids, b, s_, f, m, BT_DEVICE, bt_feat)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        barrier()
        if profile:
            ctx.profile_enable(True)
        launches0 = ctx.launch_count
        for k in range(K):
            b, s_, f, m = place_inputs(1 + W + k)
            if l2_flush:
                flush.zero_()                               # L2 flush between timed steps (untimed)
            torch.cuda.synchronize()
            ev[k][0].record(cu_stream)
            ctx.update_streams_raw(sids, b, s_, f, m, BT_DEVICE, bt_feat)
            ev[k][1].record(cu_stream)
        launches = ctx.launch_count - launches0
        barrier()
        return [a.elapsed_time(b_) for a, b_ in ev], launches

    def run_long_pass(n_steps=200):
        """Per-step statistics over many steady-state steps (min / median / p99): the timed frames are
        played forward and backward (a ping-pong keeps every track's motion continuous, so the scene stays in
        steady state for as long as we like); no L2 flush, the K distinct 8 MB frames are larger than L2."""
        ctx.tracker_reset(cfg)
        for i in range(0, 1 + W):
            b, s_, f, m = place_inputs(i)
            ctx.update_streams_raw(sids, b, s_, f, m, BT_DEVICE, bt_feat)
        order = list(range(1 + W, 1 + W + K)) + list(range(W + K - 1, W + 1, -1))
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        barrier()
        for k in range(n_steps):
            b, s_, f, m = place_inputs(order[k % len(order)])
            torch.cuda.synchronize()
            ev[k][0].record(cu_stream)
            ctx.update_streams_raw(sids, b, s_, f, m, BT_DEVICE, bt_feat)
            ev[k][1].record(cu_stream)
        barrier()
        live = sum(int(len(ctx.get_tracks(0, stream=k)["ids"])) for k in range(S))
        return sorted(a.elapsed_time(b_) for a, b_ in ev), live

    def run_tight_pass():
        """Device-resident inputs in a tight loop, the way a pipeline that owns the GPU calls the library:
        bt_submit_streams(frame k+1, device pointers) then bt_step_streams(frame k), nothing else between the calls
        (the flushed `value` protocol runs torch ops -- memset, copies, synchronisations -- between two steps, which
        leaves the host side of the next step cold).  The next frame's device-to-device placement into the ctx's
        double-buffered association buffers runs on the copy stream under the current step; K distinct 8 MB frames
        are larger than L2.  Wall clock over the K steps."""
        ctx.tracker_reset(cfg)
        for i in range(0, 1 + W):
            b, s_, f, m = ptrs(i, "d")
            ctx.update_streams_raw(sids, b, s_, f, m, BT_DEVICE, bt_feat)
        first = 1 + W
        b, s_, f, m = ptrs(first, "d")
        ctx.submit_streams_raw(sids, b, s_, f, m, BT_DEVICE, bt_feat)
        nxt = [ptrs(first + k + 1, "d") for k in range(K - 1)]
        barrier()
        t_begin = time.perf_counter()
        for k in range(K):
            if k + 1 < K:
                b, s_, f, m = nxt[k]
                ctx.submit_streams_raw(sids, b, s_, f, m, BT_DEVICE, bt_feat)
            ctx.step_streams_raw(sids)
        ctx.sync()
        total_ms = 1e3 * (time.perf_counter() - t_begin)
        barrier()
        return total_ms

    def run_e2e_pass(pipelined):
        """Pinned host inputs: every step's host->device copy and the read-back of its tracks are inside the
        timed region.  pipelined: bt_submit_streams(frame k+1) is issued before bt_step_streams(frame k), so
        the copy of the next frame runs under the current frame's step (both inside the timed region)."""
        ctx.tracker_reset(cfg)
        for i in range(0, 1 + W):
            b, s_, f, m = ptrs(i, "h")
            ctx.update_streams_raw(sids, b, s_, f, m, BT_HOST, bt_feat)
        barrier()
        d2h = 0
        first = 1 + W
        t_begin = time.perf_counter()
        if pipelined:
            b, s_, f, m = ptrs(first, "h")
            ctx.submit_streams_raw(sids, b, s_, f, m, BT_HOST, bt_feat)
        for k in range(K):
            i = first + k
            if pipelined:
                if k + 1 < K:
                    b, s_, f, m = ptrs(i + 1, "h")
                    ctx.submit_streams_raw(sids, b, s_, f, m, BT_HOST, bt_feat)
                ctx.step_streams_raw(sids)
            else:
                b, s_, f, m = ptrs(i, "h")
                ctx.update_streams_raw(sids, b, s_, f, m, BT_HOST, bt_feat)
            d2h = read_back()
        ctx.sync()
        total_ms = 1e3 * (time.perf_counter() - t_begin)
        barrier()
        return total_ms, d2h

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    run_value_pass()                                         # untimed: first-launch / module-load costs
    dev_ms, launches = run_value_pass()                      # `value`
    # the other admissible protocol: no flush, every step reads a detection frame it has never touched
    dev_ms_warm, _ = run_value_pass(l2_flush=False)
    long_ms, long_live = run_long_pass(200) if K >= 2 else ([0.0], 0)
    run_value_pass(profile=True)                             # same steps again with per-kernel CUDA events
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    assoc_replay_ms = ctx.profile_replay_assoc(50) if (n > 0) else 0.0   # the frame's association kernel, 50 back-to-back launches
    n_live = sum(int(len(ctx.get_tracks(0, stream=k)["ids"])) for k in range(S))
    run_tight_pass()
    tight_ms = run_tight_pass()
    # ---- end-to-end passes ----
    run_e2e_pass(True)
    e2e_ms, d2h_bytes = run_e2e_pass(True)
    e2e_plain_ms, _ = run_e2e_pass(False)
    # the timed regions are a few ms, shorter than nvidia-smi's sampling period: keep the same
    # device-resident step running for ~0.6 s so the clock record covers this exact load
    t_end = time.perf_counter() + 0.6
    while rank == 0 and time.perf_counter() < t_end:
        for i in range(1 + W, 1 + W + K):
            b, s_, f, m = ptrs(i, "d")
            ctx.update_streams_raw(sids, b, s_, f, m, BT_DEVICE, bt_feat)
    clocks = sampler.stop() if rank == 0 else None

    total_dev_ms, total_e2e_ms, total_warm_ms, total_plain_ms, total_tight_ms = max_over_ranks(
        [sum(dev_ms), e2e_ms, sum(dev_ms_warm), e2e_plain_ms, tight_ms], device="cuda")
    value = aggregate_throughput(n * S, world, K, total_dev_ms)
    e2e_value = aggregate_throughput(n * S, world, K, total_e2e_ms)
    feat_bytes = 2 if f16 else 4
    h2d_bytes = S * (n * (16 + 4) + (n * D * feat_bytes if reid else 0))

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        dev_sorted = sorted(dev_ms)
        assoc_ms, assoc_n = prof["assoc"]
        assoc_in_step_ms = assoc_ms / max(1, assoc_n)
        assoc_avg_ms = assoc_replay_ms
        if reid and D >= 512:
            flops = 2.0 * n_live * n * D
            peak = peaks.get("bf16_tflops", 1590.0)
            traffic, traffic_src = ncu_traffic("assoc_tc_kernel", (n_live // max(1, S), n, D, S))
            roof = {"kernel": "assoc_tc_kernel (fused ReID GEMM + IoU + cost fusion + candidate emission)",
                    "bound": "tensor", "achieved": flops / (assoc_avg_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                    "peak_source": ("MEASURED_PEAKS.json bf16_tflops (burst; fp16 runs at the same tensor rate)"
                                    if peaks else "fallback 1590 (B200_PROFILING.md)"),
                    "algorithmic_flops": flops, "avg_launch_ms": assoc_avg_ms,
                    "how": "one CUDA-event pair around 50 back-to-back launches of the last frame's kernel on the ctx "
                           "stream (bt_profile_replay_assoc), divided by 50; launched as in the step (programmatic "
                           "dependent launch: a launch's main loop starts while its predecessor drains, only its epilogue "
                           "waits for it)",
                    "in_step_event_ms": assoc_in_step_ms,
                    "in_step_note": "events bracketing the same launch inside the step also count host enqueue gaps and "
                                    "the part of the prep kernel the launch overlaps",
                    "traffic": traffic, "traffic_source": traffic_src,
                    "traffic_unit": "bytes/launch (algorithmic: the two fp16 operands once = %d)" % (2 * (n_live + n * S) * D)}
        else:
            nbytes = 32.0 * (n_live + n * S) + 0.0      # boxes in, candidate edges out (sparse)
            peak = peaks.get("hbm_gbs", 6650.0)
            roof = {"kernel": "assoc_simt_kernel (IoU-only / small-feature candidate emission)", "bound": "hbm",
                    "achieved": nbytes / (assoc_avg_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                    "algorithmic_bytes": nbytes, "avg_launch_ms": assoc_avg_ms, "in_step_event_ms": assoc_in_step_ms,
                    "traffic": None}
        roof["frac"] = roof["achieved"] / roof["peak"]
        segs = {k: (v[0] / max(1, v[1])) for k, v in prof.items()}

        cpu = None
        if world == 1 and not args.no_cpu:
            _all_host_threads()
            cf = make_frames(wl, 1 + 1 + args.cpu_frames, seed=1234)
            t_vec = cpu_frame_times(wl, cf, "vectorized")[1:]
            n_ref = max(1, args.cpu_frames // 2)
            t_ref = cpu_frame_times(wl, cf[: 1 + 1 + n_ref], "faithful")[1:]       # full frames, no extrapolation
            v_ref = n / float(np.mean(t_ref))
            v_vec = n / float(np.mean(t_vec))
            cpu = {"value": v_ref, "unit": "tracks/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": (f"{len(t_ref)} full frames of the same workload, oracle port in the reference's loop structure "
                              f"(pure-Python IoU over all {n}x{n} pairs, per-match Kalman update), "
                              f"{sum(t_ref):.1f} s of CPU work; nothing extrapolated"),
                   "vectorized_numpy": {"value": v_vec, "unit": "tracks/s",
                                        "sample": f"{len(t_vec)} frames, broadcast IoU + batched Kalman update"},
                   "ratios": {"e2e_over_reference_structure": e2e_value / v_ref, "e2e_over_vectorized_numpy": e2e_value / v_vec,
                              "value_over_reference_structure": value / v_ref, "value_over_vectorized_numpy": value / v_vec}}
        line = {
            "metric": METRIC, "value": value, "unit": "tracks/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("f64 Kalman/IoU/LAP, fp16 tcgen05 ReID similarity (fp32 accumulate), exact fp32/f64 re-costing of contested edges"
                      if reid else "f64"),
            "data": "synthetic",
            "config": config_of(wl, S),
            "feat_dtype": args.feat_dtype,
            "live_tracks_end": n_live,
            "step_ms": {"min": dev_sorted[0], "median": dev_sorted[len(dev_sorted) // 2], "max": dev_sorted[-1]},
            "e2e": {"value": e2e_value, "unit": "tracks/s", "ms_per_step": total_e2e_ms / K,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(d2h_bytes),
                    "how": "pinned host inputs: bt_submit_streams(frame k+1) then bt_step_streams(frame k) + bt_get_tracks; "
                           "every step's H2D copy and result read-back are inside the timed region (wall clock over the K steps)",
                    "unpipelined_ms_per_step": total_plain_ms / K,
                    "unpipelined_how": "bt_update_streams on the same pinned buffers (copy, then step), wall clock"},
            "value_inputs_larger_than_l2": {
                "value": aggregate_throughput(n * S, world, K, total_warm_ms), "unit": "tracks/s",
                "ms_per_step": total_warm_ms / K,
                "note": "same K steps without the L2 flush"},
            "value_tight_loop": {
                "value": aggregate_throughput(n * S, world, K, total_tight_ms), "unit": "tracks/s",
                "ms_per_step": total_tight_ms / K,
                "note": "device-resident inputs, bt_submit_streams(frame k+1) + bt_step_streams(frame k) back to back with "
                        "nothing between the calls (wall clock over the K steps, max over ranks); the next frame's "
                        "placement into the association buffers overlaps the step; distinct frames larger than L2"},
            "steady_200": {"steps": len(long_ms), "min_ms": long_ms[0], "median_ms": long_ms[len(long_ms) // 2],
                           "p99_ms": long_ms[min(len(long_ms) - 1, int(0.99 * len(long_ms)))], "mean_ms": sum(long_ms) / len(long_ms),
                           "live_tracks_end": long_live,
                           "note": "rank 0: 200 consecutive steady-state steps (the timed frames played forward and backward), "
                                   "per-step CUDA events, no L2 flush (distinct frames larger than L2)"},
            "gpu_launches": int(launches),
            "segments_ms": segs,
            "roofline": roof,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "host_placement": ({"cores": [host_cores[0], host_cores[-1]], "n_cores": len(host_cores),
                                "gpu_local_cpulist_known": bool(topo[local])} if host_cores else None),
        }
    ctx.close()
    c5 = None
    if not args.no_c5 and args.workload == "c3":
        del data, flush
        torch.cuda.empty_cache()
        c5 = run_c5(args, rank, world, local, dev, K, W)
    if rank == 0:
        line["c5"] = c5
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# BASELINE config 5: 32 independent video streams x 1000 tracks, block-partitioned over the ranks,
# per-stream NCCL scatter of the detector outputs from rank 0 and gather of ids + boxes back
# ------------------------------------------------------------------------------------------------
class _RawCuda:
    """A device buffer owned by the tracker ctx, exposed to torch without a copy."""
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def run_c5(args, rank, world, local, dev, K, W):
    import torch
    import torch.distributed as dist
    import botsort_b200 as bs
    from botsort_b200._lib import BT_DEVICE, BT_F16
    from botsort_b200.sharding import (frame_slot_ints, gather_streams, max_over_ranks, pack_frame, pack_result,
                                       result_slot_ints, scatter_streams, shard_streams, unpack_frame, unpack_result)
    from botsort_b200.synthetic import SceneConfig, SyntheticScene
    n_streams, n, D = 32, 1000, 2048
    mine = shard_streams(n_streams, world, rank)
    S = len(mine)
    cap = 1152
    K5, W5 = min(K, 10), 3
    n_frames = 1 + W5 + K5
    ctx = bs.Context(max_tracks=cap, max_dets=cap, feat_dim=D, device=local, n_streams=S)
    cfg = ctx.default_config()
    ctx.tracker_reset(cfg)
    sids = list(range(S))
    # identity banks of the streams this GPU owns ("features produced on the owning GPU": the stand-in for its
    # own TensorRT ReID engine, which writes fp16 rows straight into the association buffers)
    scene_cfg = lambda sid: SceneConfig(n_ids=n, feat_dim=D, seed=5000 + sid, emit_features=False)
    ident = [torch.from_numpy(SyntheticScene(scene_cfg(sid)).identity).to(dev) for sid in mine]
    all_scenes = [SyntheticScene(scene_cfg(sid)) for sid in range(n_streams)] if rank == 0 else None
    all_ident16 = None
    if rank == 0 and args.c5_full_scatter:
        all_ident16 = [torch.from_numpy(sc.identity).to(dev) for sc in all_scenes]
    slot_f, slot_r = frame_slot_ints(cap), result_slot_ints(cap)
    # rank 0 is the producer: every stream's detector outputs of every frame, packed into per-stream slots up front
    # (scene generation is not part of the scatter)
    stage_h = None
    if rank == 0:
        stage_h = torch.empty((n_frames, n_streams, slot_f), dtype=torch.int32).pin_memory()
        for i in range(n_frames):
            for sid, sc in enumerate(all_scenes):
                fr = sc.next_frame()
                stage_h[i, sid].copy_(torch.from_numpy(pack_frame(fr["boxes"], fr["scores"], fr["gt"], cap)))
    res_h = torch.empty((S, slot_r), dtype=torch.int32).pin_memory()
    gen = torch.Generator(device=dev)
    gen.manual_seed(77 + rank)
    views = {}

    def in_views(k):
        pb, ps, pf = ctx.input_buffers(k)
        key = (k, pb)
        if key not in views:
            views[key] = (pb, ps, pf,
                          torch.as_tensor(_RawCuda(pb, (cap, 4), "<i4"), device=dev),
                          torch.as_tensor(_RawCuda(ps, (cap,), "<f4"), device=dev),
                          torch.as_tensor(_RawCuda(pf, (cap, D), "<f2"), device=dev))
        return views[key]

    t_scatter = t_step = t_gather = t_feat = t_full = 0.0
    step_list = []
    checked = 0
    for i in range(n_frames):
        timed = i >= 1 + W5
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        # ---- scatter: rank 0 packs every stream's detector output, one NCCL scatter of per-stream slots ----
        packed = None
        if rank == 0:
            packed = stage_h[i].to(dev, non_blocking=True)
        recv = scatter_streams(packed, n_streams, slot_f, src=0, device=dev)
        ms = recv[:, 0].cpu().tolist()
        bufs = []
        gts = []
        for k in range(S):
            pb, ps, pf, vb, vs, vf = in_views(k)
            m, boxes, scores, gt = unpack_frame(recv[k], cap)
            vb[:m].copy_(boxes)
            vs[:m].copy_(scores)
            bufs.append((pb, ps, pf, m, vf))
            gts.append(gt.long())
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        # ---- features on the owning GPU (ReID stand-in, not part of the tracker's time) ----
        for k in range(S):
            m, vf = bufs[k][3], bufs[k][4]
            f = ident[k][gts[k]] + 0.005 * torch.randn((m, D), device=dev, generator=gen)
            f = f / f.norm(dim=1, keepdim=True)
            vf[:m].copy_(f.half())
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        # ---- the tracker: all local streams in ONE batched frame step ----
        ctx.update_streams_raw(sids, [b[0] for b in bufs], [b[1] for b in bufs], [b[2] for b in bufs],
                               [b[3] for b in bufs], BT_DEVICE, BT_F16)
        t3 = time.perf_counter()
        # ---- gather: ids + boxes of every stream back to rank 0 ----
        for k in range(S):
            tr = ctx.get_tracks(0, stream=k)
            res_h[k].copy_(torch.from_numpy(pack_result(tr["ids"], tr["tlbr"], cap)))
        allres = gather_streams(res_h.to(dev, non_blocking=True), n_streams, slot_r, dst=0, device=dev)
        if rank == 0:
            host = allres.cpu().numpy()
            if i == n_frames - 1:
                for sid in range(n_streams):
                    ids, tlbr = unpack_result(host[sid], cap)
                    checked += int(len(ids))
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        if timed:
            t_scatter += t1 - t0; t_feat += t2 - t1; t_step += t3 - t2; t_gather += t4 - t3
            step_list.append(t3 - t2)
        # ---- mode 2: what scattering the FEATURES from one rank would cost (NVLink), timed on its own ----
        if timed and args.c5_full_scatter and world > 1:
            parts = None
            if rank == 0:
                parts = [torch.empty((len(shard_streams(n_streams, world, r)), cap, D), dtype=torch.float16, device=dev) for r in range(world)]
            mine_f = torch.empty((S, cap, D), dtype=torch.float16, device=dev)
            dist.barrier(); torch.cuda.synchronize()
            tf0 = time.perf_counter()
            dist.scatter(mine_f, parts, src=0)
            torch.cuda.synchronize()
            t_full += time.perf_counter() - tf0
    step_med = float(np.median(step_list)) if step_list else 0.0
    totals = max_over_ranks([t_scatter, t_step, t_gather, t_feat, t_full, step_med * K5], device="cuda")
    live = sum(int(len(ctx.get_tracks(0, stream=k)["ids"])) for k in range(S))
    ctx.close()
    if rank != 0:
        return None
    sc_ms, st_ms, ga_ms, fe_ms, fu_ms, st_med_ms = (1e3 * v / K5 for v in totals)
    return {
        "workload": "32 independent synthetic video streams x 1000 tracks x 1000 dets x 2048-d, block-partitioned over the GPUs "
                    f"({n_streams // world} streams per GPU, ONE batched frame step per GPU and frame)",
        "n_gpus": world, "streams": n_streams, "streams_per_gpu": n_streams // world, "steps": K5, "warmup": W5,
        "value": n_streams * n / (st_ms / 1e3), "unit": "tracks/s",
        "step_ms": st_ms,
        "step_ms_median": st_med_ms,
        "value_median_step": n_streams * n / (st_med_ms / 1e3) if st_med_ms > 0 else None,
        "median_note": "step_ms = mean over the timed frames (max over ranks), step_ms_median = each rank's median (max over "
                       "ranks): the step is host-bound wall clock, a single descheduling of a rank's thread shows in the mean",
        "value_incl_scatter_gather": n_streams * n / ((sc_ms + st_ms + ga_ms) / 1e3),
        "scatter_ms": sc_ms, "gather_ms": ga_ms,
        "scatter_bytes_per_frame": n_streams * slot_f * 4, "gather_bytes_per_frame": n_streams * slot_r * 4,
        "feature_production_ms_excluded": fe_ms,
        "full_feature_scatter_ms": (fu_ms if (args.c5_full_scatter and world > 1) else None),
        "full_feature_scatter_bytes_per_frame": n_streams * cap * D * 2,
        "mode": "features produced on the owning GPU (fp16 rows written into the association buffers in place); NCCL carries "
                "boxes / scores / identities out and ids / boxes back; full_feature_scatter_ms = the NVLink cost of scattering "
                "the feature rows from rank 0 instead (SURVEY 8(e) mode 2), timed separately",
        "how": "wall clock per phase with device synchronisation on both sides, max over ranks; frame step = bt_update_streams "
               "(device-resident inputs, results on the host when it returns)",
        "tracks_gathered_last_frame": checked, "live_tracks_rank0": live,
    }


# ------------------------------------------------------------------------------------------------
# BASELINE config 4: raw YOLOX head + camera frame -> bt_detect_stage -> stub encoder -> bt_update_streams
# ------------------------------------------------------------------------------------------------
C4_POOL = 16          # the stub encoder: 16x16 average pooling of the crop tensor, then a fixed random projection


def c4_projection(D):
    return (np.random.default_rng(7).standard_normal((3 * (256 // C4_POOL) * (128 // C4_POOL), D)) / 8.0).astype(np.float32)


def c4_frames(wl, count, seed):
    from botsort_b200.synthetic import DetectorScene
    sc = DetectorScene(k=wl["persons"], h=wl["img_h"], w=wl["img_w"], seed=seed)
    return [sc.next_frame() for _ in range(count)]


def c4_cpu_times(wl, frames):
    """The same chain on the host cores: oracle decode + NMS + _postprocess, cv2-exact crop + normalise, the stub
    encoder in NumPy, the oracle tracker in the reference's loop structure.  First frame untimed (births)."""
    from oracle import detector_np as Dn
    from oracle import oracle_np as O
    proj = c4_projection(wl["feat_dim"])
    trk = O.OracleBoTSORT(mode="faithful", lap_solver="jv", use_features=True)
    times, parts, n_live = [], [], 0
    for k, (frame, raw) in enumerate(frames):
        t0 = time.perf_counter()
        det = Dn.yolox_postprocess(raw, img_h=wl["img_h"], img_w=wl["img_w"])
        body = det[det[:, 0] == 0]
        b_int = body[:, 2:6].astype(np.int32)
        t1 = time.perf_counter()
        crops = Dn.crop_preprocess(frame, b_int)
        t2 = time.perf_counter()
        pooled = crops.reshape(len(b_int), 3, 256 // C4_POOL, C4_POOL, 128 // C4_POOL, C4_POOL).mean(axis=(3, 5))
        f = pooled.reshape(len(b_int), -1).astype(np.float32) @ proj
        f = (f / np.maximum(np.linalg.norm(f, axis=1, keepdims=True), 1e-12)).astype(np.float16).astype(np.float32)
        t3 = time.perf_counter()
        trk.update_arrays(b_int, body[:, 1].astype(np.float32), f)
        t4 = time.perf_counter()
        if k > 0:
            times.append(t4 - t0)
            parts.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
        n_live = len(trk.snapshot()["tracked"]["ids"])
    return times, np.mean(np.array(parts), axis=0) if parts else np.zeros(4), n_live


def c4_config(wl):
    c = config_of(wl, 1)
    c.update({"tracks": wl["persons"], "dets": "<= 50 bodies / frame (max_per_class)", "anchors": 6300, "classes": 4,
              "crop": "3x256x128 float32 per body", "encoder": "stub (16x16 average pool + fixed 384x2048 projection, L2 "
              "normalise, fp16) -- FastReID inference itself is out of scope (north_star: stays on TensorRT)"})
    return c


def run_c4_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = WORKLOADS["c4"]
    _all_host_threads()
    K, W = args.steps, args.warmup
    frames = c4_frames(wl, 1 + W + K, seed=1234)
    t0 = time.perf_counter()
    t_all, parts, n_live = c4_cpu_times(wl, frames)
    t = t_all[W:]
    ms = 1e3 * float(np.mean(t))
    value = n_live / (ms / 1e3)
    sample = (f"{len(t)} timed frames (+{W} warm-up) of the full chain on the host: decode+NMS "
              f"{1e3 * parts[0]:.2f} ms, crop+normalise {1e3 * parts[1]:.2f} ms, stub encoder {1e3 * parts[2]:.2f} ms, "
              f"tracker update {1e3 * parts[3]:.2f} ms per frame")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "tracks/s", "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": c4_config(wl), "live_tracks_end": n_live, "frames_per_s": 1e3 / ms,
        "run_wall_s": time.perf_counter() - t0,
        "cpu_baseline": {"value": value, "unit": "tracks/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "tracks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_c4(args):
    import ctypes as C
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F

    import botsort_b200 as bs
    from botsort_b200._lib import BT_DEVICE, BT_F16
    from botsort_b200.sharding import aggregate_throughput, max_over_ranks

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS["c4"]
    D, H, Wd = wl["feat_dim"], wl["img_h"], wl["img_w"]
    K, W = args.steps, max(args.warmup, 3)
    ctx = bs.Context(max_tracks=256, max_dets=256, feat_dim=D, device=local)
    ycfg = bs.BtYoloxConfig()
    ctx.lib.bt_default_yolox_config(C.byref(ycfg))
    mb = int(ycfg.max_per_class)
    cu_stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    n_frames = 1 + W + K
    frames = c4_frames(wl, n_frames, seed=1234 + rank)
    fh = [torch.from_numpy(f).pin_memory() for f, _ in frames]
    rh = [torch.from_numpy(r).pin_memory() for _, r in frames]
    fd = [x.to(dev) for x in fh]
    rd = [x.to(dev) for x in rh]
    proj = torch.from_numpy(c4_projection(D)).to(dev)
    crops = torch.zeros((mb, 3, 256, 128), dtype=torch.float32, device=dev)
    det_out = torch.zeros((256, 6), dtype=torch.float64, device=dev)
    det_cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    stage_f = torch.empty((H, Wd, 3), dtype=torch.uint8, device=dev)
    stage_r = torch.empty_like(rd[0])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def chain(d_raw, d_frame):
        """One frame, everything enqueued on the ctx stream, nothing visits the host between the stages."""
        pb, ps, pf = ctx.input_buffers(0)
        ctx.detect_stage(d_raw.data_ptr(), d_frame.data_ptr(), H, Wd, crops.data_ptr(), ycfg)
        with torch.cuda.stream(cu_stream):                   # the stub encoder (library ops; stands in for TensorRT)
            f32 = F.avg_pool2d(crops, C4_POOL).reshape(mb, -1) @ proj
            f32 = f32 / f32.norm(dim=1, keepdim=True).clamp_min(1e-12)
            torch.as_tensor(_RawCuda(pf, (mb, D), "<f2"), device=dev).copy_(f32.half())
        ctx.update_streams_raw([0], [pb], [ps], [pf], [mb], BT_DEVICE, BT_F16)

    def run_value_pass():
        ctx.tracker_reset()
        for i in range(1 + W):
            chain(rd[i], fd[i])
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        barrier()
        l0 = ctx.launch_count
        for k in range(K):
            flush.zero_()
            torch.cuda.synchronize()
            ev[k][0].record(cu_stream)
            chain(rd[1 + W + k], fd[1 + W + k])
            ev[k][1].record(cu_stream)
        launches = ctx.launch_count - l0
        barrier()
        return [a.elapsed_time(b) for a, b in ev], launches

    def run_e2e_pass():
        ctx.tracker_reset()
        for i in range(1 + W):
            chain(rd[i], fd[i])
        barrier()
        t0 = time.perf_counter()
        nbytes = 0
        for k in range(K):
            with torch.cuda.stream(cu_stream):
                stage_f.copy_(fh[1 + W + k], non_blocking=True)
                stage_r.copy_(rh[1 + W + k], non_blocking=True)
            chain(stage_r, stage_f)
            res = ctx.get_tracks(0)
            nbytes = res["tlbr"].nbytes + res["ids"].nbytes
        ctx.sync()
        ms = 1e3 * (time.perf_counter() - t0)
        barrier()
        return ms, nbytes

    def time_kernel(fn):
        """K stand-alone launches of one detector-side kernel on the ctx stream.  Before each one the L2 is flushed ON
        THE SAME STREAM (2 x 256 MiB memset) and the launch is enqueued behind it without a host synchronisation, so
        the event pair brackets the kernel alone and not the host's enqueue latency."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for k in range(K):
            with torch.cuda.stream(cu_stream):
                flush.zero_()
                flush.zero_()
            ev[k][0].record(cu_stream)
            fn(k)
            ev[k][1].record(cu_stream)
            torch.cuda.synchronize()
        t = sorted(a.elapsed_time(b) for a, b in ev)
        return float(np.mean(t)), t[len(t) // 2]

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    run_value_pass()
    dev_ms, launches = run_value_pass()
    n_live = int(len(ctx.get_tracks(0)["ids"]))
    n_body = int((det_out[:, 0] == 0).sum().item()) if False else None
    run_e2e_pass()
    e2e_ms, d2h = run_e2e_pass()
    # ---- per-kernel durations (stand-alone device-resident launches of the same frames) ----
    boxes_dev = torch.zeros((mb, 4), dtype=torch.int32, device=dev)

    def yolox_only(k):
        ctx._check(ctx.lib.bt_yolox_postprocess(ctx.h, C.c_void_p(rd[1 + W + k].data_ptr()), C.byref(ycfg),
                                                C.c_void_p(det_out.data_ptr()), 256, C.c_void_p(det_cnt.data_ptr()), BT_DEVICE))
    yolox_ms, yolox_med = time_kernel(yolox_only)
    torch.cuda.synchronize()
    det = det_out.cpu().numpy()[: int(det_cnt.cpu()[0])]
    body = det[det[:, 0] == 0][:, 2:6].astype(np.int32)
    nb = len(body)
    boxes_dev[:nb] = torch.from_numpy(body).to(dev)

    def crop_only(k):
        ctx._check(ctx.lib.bt_reid_crop_gather(ctx.h, C.c_void_p(fd[1 + W + k].data_ptr()), H, Wd, C.c_void_p(boxes_dev.data_ptr()),
                                               nb, 256, 128, C.c_void_p(crops.data_ptr()), BT_DEVICE))
    crop_ms, crop_med = time_kernel(crop_only)
    t_end = time.perf_counter() + 0.6
    while rank == 0 and time.perf_counter() < t_end:
        for i in range(1 + W, 1 + W + K):
            chain(rd[i], fd[i])
    ctx.sync()
    clocks = sampler.stop() if rank == 0 else None
    total_dev_ms, total_e2e_ms = max_over_ranks([sum(dev_ms), e2e_ms], device="cuda")
    value = aggregate_throughput(n_live, world, K, total_dev_ms)
    e2e_value = aggregate_throughput(n_live, world, K, total_e2e_ms)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = peaks.get("hbm_gbs", 6650.0)
        crop_bytes = float(nb) * 3 * 256 * 128 * 4
        yolo_bytes = float(rd[0].numel() * 4)
        crop_traffic, crop_src = ncu_traffic("reid_crop_kernel", (nb, 256, 128))
        yolo_traffic, yolo_src = ncu_traffic("yolox_post_kernel", (6300, 4))
        roof = {"kernel": "reid_crop_kernel (crop + INTER_LINEAR resize + BGR->RGB + normalise, written once)", "bound": "hbm",
                "achieved": crop_bytes / (crop_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                "algorithmic_bytes": crop_bytes, "bytes_per_det": 3 * 256 * 128 * 4, "dets": nb, "avg_launch_ms": crop_ms,
                "median_launch_ms": crop_med, "traffic": crop_traffic, "traffic_source": crop_src,
                "how": "CUDA events on the ctx stream around K stand-alone device-resident launches, each enqueued behind an "
                       "L2 flush on the same stream (no host latency inside the bracket)"}
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof_y = {"kernel": "yolox_post (decode + per-class NMS + _postprocess)", "bound": "hbm",
                  "achieved": yolo_bytes / (yolox_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                  "algorithmic_bytes": yolo_bytes, "avg_launch_ms": yolox_ms, "median_launch_ms": yolox_med,
                  "traffic": yolo_traffic, "traffic_source": yolo_src,
                  "note": "226.8 KB per frame: a latency chain (decode -> sort -> greedy NMS), not a bandwidth problem; "
                          "the duration is what matters"}
        roof_y["frac"] = roof_y["achieved"] / roof_y["peak"]
        cpu = None
        if world == 1 and not args.no_cpu:
            _all_host_threads()
            t_cpu, parts, live_cpu = c4_cpu_times(wl, frames[: 2 + max(2, args.cpu_frames)])
            v = live_cpu / float(np.mean(t_cpu))
            cpu = {"value": v, "unit": "tracks/s", "cores": os.cpu_count(), "kind": "port",
                   "sample": (f"{len(t_cpu)} full frames of the same chain on the host: decode+NMS {1e3 * parts[0]:.2f} ms, "
                              f"crop+normalise {1e3 * parts[1]:.2f} ms, stub encoder {1e3 * parts[2]:.2f} ms, tracker "
                              f"{1e3 * parts[3]:.2f} ms per frame"),
                   "ratios": {"e2e_over_cpu": e2e_value / v, "value_over_cpu": value / v}}
        ds = sorted(dev_ms)
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "tracks/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 decode/NMS, u8 fixed-point resize, fp16 features, f64 Kalman/IoU/LAP", "data": "synthetic",
            "config": c4_config(wl), "live_tracks_end": n_live, "bodies_per_frame": nb,
            "frames_per_s": 1e3 * K * world / total_dev_ms,
            "step_ms": {"min": ds[0], "median": ds[len(ds) // 2], "max": ds[-1]},
            "e2e": {"value": e2e_value, "unit": "tracks/s", "ms_per_step": total_e2e_ms / K,
                    "h2d_bytes_per_step": int(fh[0].numel() + rh[0].numel() * 4), "d2h_bytes_per_step": int(d2h),
                    "how": "pinned host frame + raw head copied in on the ctx stream, the device chain, bt_get_tracks back; "
                           "wall clock over the K steps"},
            "gpu_launches": int(launches),
            "gpu_launches_note": "this library's kernels only (the stub encoder's torch ops are not counted)",
            "roofline": roof, "rooflines": [roof, roof_y], "cpu_baseline": cpu, "clocks": clocks}))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    # torch's caching allocator still holds blocks that were used on the ctx's stream (the stub encoder ran there):
    # destroying that stream first makes torch's own teardown abort.  The process is done: leave without either teardown.
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


_CUDART = None


def _cudart():
    """libcudart through ctypes (plumbing for placing synthetic inputs at raw device addresses)."""
    global _CUDART
    if _CUDART is None:
        import ctypes
        import torch
        lib = None
        for name in ("libcudart.so.12", "libcudart.so"):
            try:
                lib = ctypes.CDLL(name)
                break
            except OSError:
                continue
        if lib is None:
            import glob
            cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + \
                glob.glob("/usr/local/cuda/lib64/libcudart.so*")
            lib = ctypes.CDLL(cands[0])
        lib.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        lib.cudaMemcpy.restype = ctypes.c_int
        _CUDART = lib
    return _CUDART


def ncu_traffic(kernel, shape):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture of this workload
    (profiles/r02_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep), or (None, why)."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        rec = json.load(open(path))
        for e in rec.get("kernels", []):
            if e["kernel"] == kernel and tuple(e["shape"]) == tuple(shape):
                return float(e["dram_bytes_per_launch"]), f"profiles/r02_traffic.json ({e.get('capture', 'ncu --set full')})"
        return None, "no ncu capture of this shape under profiles/"
    except Exception:
        return None, "profiles/r02_traffic.json absent"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-frames", type=int, default=4, help="frames of the cpu_baseline leg")
    ap.add_argument("--feat-dtype", default="f16", choices=["f16", "f32"],
                    help="dtype of the ReID feature rows fed to the tracker (fp16 = the reference's TensorRT engine precision)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-c5", action="store_true", help="skip the BASELINE config-5 sub-record (32 streams, NCCL scatter / gather)")
    ap.add_argument("--no-c5-full-scatter", dest="c5_full_scatter", action="store_false",
                    help="skip timing the NVLink scatter of the feature rows (config 5, mode 2)")
    args = ap.parse_args()
    if args.workload == "c4":
        (run_c4_reference if args.impl == "reference" else run_c4)(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
