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
    from botsort_b200._lib import BT_DEVICE, BT_HOST

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
    if local_world > 1 and hasattr(os, "sched_setaffinity"):
        # one rank per GPU on one host: give every rank its own slice of the host cores so that the ranks'
        # enqueue / bookkeeping threads (they are on the frame's critical path) do not migrate onto each other
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // local_world)
        mine = cores[local * per:(local + 1) * per] or cores
        try:
            os.sched_setaffinity(0, mine)
        except OSError:
            pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    wl = WORKLOADS[args.workload]
    n, D, reid = wl["n"], wl["feat_dim"], wl["reid"]
    K, W = args.steps, max(args.warmup, 3)

    from concurrent.futures import ThreadPoolExecutor
    from botsort_b200.sharding import shard_streams
    S = int(wl.get("streams", 1))                       # independent video streams on THIS GPU
    my_streams = shard_streams(S * world, world, rank)  # global stream ids of this rank (no data-path collective)
    cap = (n + n // 8 + 255) // 128 * 128
    ctxs = [bs.Context(max_tracks=cap, max_dets=cap, feat_dim=D, device=local) for _ in my_streams]
    ctx = ctxs[0]
    cfg = ctx.default_config()
    cfg.with_reid = 1 if reid else 0
    cu_streams = [torch.cuda.ExternalStream(c.stream, device=torch.device("cuda", local)) for c in ctxs]
    pool = ThreadPoolExecutor(max_workers=S) if S > 1 else None

    # every stream has its own synthetic scene (weak scaling: each GPU tracks its own video streams)
    n_frames = 1 + W + K
    data = []
    for sid in my_streams:
        frames = make_frames(wl, n_frames, seed=1234 + sid)
        bh = [torch.from_numpy(f["boxes"]).pin_memory() for f in frames]
        sh = [torch.from_numpy(f["scores"]).pin_memory() for f in frames]
        fh = [torch.from_numpy(f["feats"]).pin_memory() for f in frames] if reid else [None] * n_frames
        data.append(dict(bh=bh, sh=sh, fh=fh, bd=[b.cuda(non_blocking=True) for b in bh],
                         sd=[x.cuda(non_blocking=True) for x in sh],
                         fd=[f.cuda(non_blocking=True) for f in fh] if reid else [None] * n_frames))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_one(k, i, loc, read_back=False, events=None):
        d, c = data[k], ctxs[k]
        b, s, f = (d["bd"], d["sd"], d["fd"]) if loc == BT_DEVICE else (d["bh"], d["sh"], d["fh"])
        if events is not None:
            events[0].record(cu_streams[k])
        c.update_arrays_raw(b[i].data_ptr(), s[i].data_ptr(), f[i].data_ptr() if reid else 0, b[i].shape[0], loc)
        nbytes = 0
        if read_back:
            res = c.get_tracks(0)                       # ids + boxes of the returned list, on the host
            nbytes = res["tlbr"].nbytes + res["ids"].nbytes
        if events is not None:
            events[1].record(cu_streams[k])
        return nbytes

    def step(i, loc, read_back=False, events=None):
        """One frame for every stream of this rank (threads: ctypes releases the GIL, the ctx streams overlap)."""
        if pool is None:
            return step_one(0, i, loc, read_back, None if events is None else events[0])
        futs = [pool.submit(step_one, k, i, loc, read_back, None if events is None else events[k]) for k in range(S)]
        return sum(f.result() for f in futs)

    def run_pass(loc, read_back, profile=False, l2_flush=True):
        """frame 0 = births (untimed), W warm-up frames, then K timed frames.  Returns per-step device
        ms (CUDA events on the ctx stream), per-step wall ms, matched-track counts."""
        for c in ctxs:
            c.tracker_reset(cfg)
        step(0, loc)
        for i in range(1, 1 + W):
            step(i, loc)
        ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(S)]
              for _ in range(K)]
        wall = []
        d2h = 0
        barrier()
        if profile:
            ctx.profile_enable(True)        # segment events only around the K timed steps (first stream)
        launches0 = sum(c.launch_count for c in ctxs)
        for k in range(K):
            if l2_flush:
                flush.zero_()                               # L2 flush between timed steps (untimed)
            torch.cuda.synchronize()
            i = 1 + W + k
            t0 = time.perf_counter()
            d2h = step(i, loc, read_back, ev[k])
            for c in ctxs:
                c.sync()
            wall.append(1e3 * (time.perf_counter() - t0))
        launches = sum(c.launch_count for c in ctxs) - launches0
        barrier()
        # a step ends when its slowest stream ends (the streams of a step start together)
        dev = [max(a.elapsed_time(b) for a, b in evk) for evk in ev]
        return dev, wall, launches, d2h

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-resident pass (value) with segment profiling ----
    run_pass(BT_DEVICE, read_back=False)                     # untimed: first-launch / module-load costs
    dev_ms, _, launches, _ = run_pass(BT_DEVICE, read_back=False)   # `value`: no event bracketing inside
    # the other admissible protocol: no flush, every step reads a detection frame it has never touched (the
    # 1+W+K frames together exceed L2) while the tracker's own state stays as warm as it is in a running stream
    dev_ms_warm, _, _, _ = run_pass(BT_DEVICE, read_back=False, l2_flush=False)
    run_pass(BT_DEVICE, read_back=False, profile=True)       # same steps again with per-kernel CUDA events
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    assoc_replay_ms = ctx.profile_replay_assoc(50)   # the frame's association kernel, 50 back-to-back launches
    info_tracks = ctx.get_tracks(0)
    n_live = int(len(info_tracks["ids"]))
    # ---- end-to-end pass: pinned host inputs, H2D inside, result read back to the host ----
    _, e2e_wall, _, d2h_bytes = run_pass(BT_HOST, read_back=True)
    # the timed regions are a few ms, shorter than nvidia-smi's sampling period: keep the same
    # device-resident step running for ~0.6 s so the clock record covers this exact load
    t_end = time.perf_counter() + 0.6
    while rank == 0 and time.perf_counter() < t_end:
        for i in range(1 + W, 1 + W + K):
            step(i, BT_DEVICE)
    clocks = sampler.stop() if rank == 0 else None

    from botsort_b200.sharding import aggregate_throughput, max_over_ranks
    total_dev_ms, total_e2e_ms, total_warm_ms = max_over_ranks([sum(dev_ms), sum(e2e_wall), sum(dev_ms_warm)], device="cuda")
    value = aggregate_throughput(n * S, world, K, total_dev_ms)
    e2e_value = aggregate_throughput(n * S, world, K, total_e2e_ms)
    h2d_bytes = S * (n * (16 + 4) + (n * D * 4 if reid else 0))

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        assoc_ms, assoc_n = prof["assoc"]
        assoc_in_step_ms = assoc_ms / max(1, assoc_n)
        assoc_avg_ms = assoc_replay_ms
        n_rows = n_live
        if reid:
            flops = 2.0 * n_rows * n * D
            peak = peaks.get("bf16_tflops", 1590.0)
            roof = {"kernel": "assoc_tc_kernel (fused ReID GEMM + IoU + cost fusion + candidate emission)",
                    "bound": "tensor", "achieved": flops / (assoc_avg_ms * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                    "peak_source": ("MEASURED_PEAKS.json bf16_tflops (burst; fp16 runs at the same tensor rate)"
                                    if peaks else "fallback 1590 (B200_PROFILING.md)"),
                    "algorithmic_flops": flops, "avg_launch_ms": assoc_avg_ms,
                    "how": "one CUDA-event pair around 50 back-to-back launches of the last frame's kernel on the ctx "
                           "stream (bt_profile_replay_assoc), divided by 50; launched as in the step (programmatic "
                           "dependent launch: a launch's ramp overlaps its predecessor's tail)",
                    "in_step_event_ms": assoc_in_step_ms,
                    "in_step_note": "events bracketing the same launch inside the step also count host enqueue gaps",
                    # dram__bytes_read.sum + dram__bytes_write.sum of one launch, `ncu --set full` capture of this
                    # kernel on this workload (profiles/r01_ncu_assoc_tc_v3_details.csv); other workloads: not captured
                    "traffic": (16927232.0 if (n_rows, n, D) == (2000, 2000, 2048) else None),
                    "traffic_unit": "bytes/launch (algorithmic: the two fp16 operands once = %d)" % (2 * (n_rows + n) * D)}
        else:
            nbytes = 32.0 * (n_rows + n) + 0.0      # boxes in, candidate edges out (sparse)
            peak = peaks.get("hbm_gbs", 6650.0)
            roof = {"kernel": "assoc_simt_kernel (IoU-only candidate emission)", "bound": "hbm",
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
            "dtype": "f64 Kalman/IoU/LAP, fp16-in/fp32-acc tcgen05 ReID similarity" if reid else "f64",
            "data": "synthetic",
            "config": config_of(wl, S),
            "live_tracks_end": n_live,
            "e2e": {"value": e2e_value, "unit": "tracks/s", "ms_per_step": total_e2e_ms / K,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(d2h_bytes),
                    "how": "bt_update_arrays on pinned host buffers + bt_get_tracks read-back, wall clock per step"},
            "value_inputs_larger_than_l2": {
                "value": aggregate_throughput(n * S, world, K, total_warm_ms), "unit": "tracks/s",
                "ms_per_step": total_warm_ms / K,
                "note": "same K steps without the L2 flush: each step's detection frame (%.1f MB) is one of %d distinct "
                        "device-resident frames (%.0f MB in total, larger than L2) and has not been touched since its "
                        "upload; the tracker's own state stays warm as in a running stream"
                        % (h2d_bytes / S / 1e6, n_frames, n_frames * h2d_bytes / S / 1e6)},
            "gpu_launches": int(launches),
            "segments_ms": segs,
            "roofline": roof,
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line))
    for c in ctxs:
        c.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-frames", type=int, default=4, help="frames of the cpu_baseline leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
