#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the GeoFlow-SLAM hot path on B200.

Workload (BASELINE.json configs[1]): a batch of 1024 synthetic 640x480 gray frames, 1000 ORB
features (scale 1.2, 8 levels, FAST 25/7), ORB extraction of every frame followed by BF-Hamming
+ GMS matching of every consecutive pair.  One "step" = one pass over the batch.  Metric:
frames/s (whole job, all ranks).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]
  torchrun --nproc-per-node N bench.py --gpus N ...      (one rank per GPU, weak scaling)

`value`  : inputs resident in HBM, timed with CUDA events on the launching stream.
`e2e`    : the same step through the host-buffer C-ABI call (gfs_frontend_run): pinned host images
           in, every result (keypoints, descriptors, matches, inlier masks) back on the host.
`roofline`: dominant kernel, algorithmic bytes (DESIGN.md "Algorithmic bytes") / CUDA-event time.
`cpu_baseline` / `--impl reference`: the restated reference CPU path (oracle/) on the host cores.
"""
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

import numpy as np  # noqa: E402

W, H = 640, 480
ORB_CFG = dict(nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=25, minThFAST=7)
LEVELS = [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231), (257, 193), (214, 161), (179, 134)]
METRIC = "frames/sec VGA ORB(1k feats)+BF/GMS match, batch 1024 (BASELINE configs[1])"


def _gen_chunk(args):
    from geoflowslam_b200 import synth
    start, n, seed0 = args
    # frames of one 8-frame group share a scene; chunks start on group boundaries
    return synth.orb_frames(n, W, H, group=8, seed0=seed0 + start)


def make_frames(n, seed0, procs):
    """n distinct synthetic frames (numpy generator, fork pool; call before CUDA is initialised)."""
    from multiprocessing import get_context
    chunk = 8
    jobs = [(s, min(chunk, n - s), seed0) for s in range(0, n, chunk)]
    if procs <= 1 or len(jobs) == 1:
        parts = [_gen_chunk(j) for j in jobs]
    else:
        with get_context("fork").Pool(procs) as pool:
            parts = pool.map(_gen_chunk, jobs)
    return np.ascontiguousarray(np.concatenate(parts, 0))


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the restated reference path (oracle) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_frames_per_sec(frames, threads):
    """ORB extract every frame + BF/GMS every consecutive pair with `threads` worker threads (each
    worker runs the single-threaded oracle; ctypes releases the GIL).  Returns (fps, seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    O.lib()
    tl = threading.local()

    def ext(i):
        if not hasattr(tl, "o"):
            tl.o = O.OrbOracle(ORB_CFG["nfeatures"], ORB_CFG["scaleFactor"], ORB_CFG["nlevels"],
                               ORB_CFG["iniThFAST"], ORB_CFG["minThFAST"], threads=1)
        return tl.o.extract(frames[i])

    def match(p):
        (k1, d1, _), (k2, d2, _) = res[p], res[p + 1]
        idx, _ = O.bf_match(d1, d2, threads=1)
        m = np.stack([np.arange(len(idx), dtype=np.int32), idx], 1)
        return O.gms_filter(np.stack([k1["x"], k1["y"]], 1), (W, H), np.stack([k2["x"], k2["y"]], 1), (W, H), m)[1]

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        res = list(ex.map(ext, range(len(frames))))
        list(ex.map(match, range(len(frames) - 1)))
    dt = time.perf_counter() - t0
    return len(frames) / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = max(16 * cores, 64)  # ~1 s of wall per step on the host cores: a bounded sample of the 1024-frame batch
    frames = make_frames(n, 1000, min(cores, 16))
    for _ in range(args.warmup):
        cpu_frames_per_sec(frames[:max(cores, 8)], cores)
    t = []
    for _ in range(args.steps):
        fps, dt = cpu_frames_per_sec(frames, cores)
        t.append(dt)
    ms = 1e3 * sum(t) / len(t)
    value = n / (ms / 1e3)
    sample = "%d frames + %d pairs per step, %d steps" % (n, n - 1, args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "configs[1]: ORB(1000 feats, 1.2, 8 levels, FAST 25/7) + BF-Hamming + GMS, 640x480",
                       "sample": sample},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(stage, n_frames, n_kp, n_cand):
    """Algorithmic bytes one launch group of `stage` moves for n_frames."""
    px = [w * h for (w, h) in LEVELS]
    per_frame = {
        "pyramid": sum(px[:-1]) + sum(px[1:]),           # read levels 0..6, write levels 1..7
        "fast_cells": sum(px) + 4 * n_cand,              # read every level once, write packed candidates
        "octree": 4 * n_cand + 4 * n_kp,                 # read candidates, write selected keys
        "blur": 0,                                       # fused into orient_desc (never materialised)
        "orient_desc": n_kp * (43 * 43 + 24 + 32 + 4),   # 43x43 patch + keypoint + descriptor + key
        "pack_lapping": 0,
        "bf_hamming": 2 * n_kp * 32 + n_kp * 8,          # both descriptor sets + (idx, dist)
        "gms": 2 * n_kp * 8 + n_kp * 4 + n_kp,           # both point sets + train idx + mask
    }
    return per_frame[stage] * n_frames


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B = args.batch
    cores = os.cpu_count() or 1
    # synthetic input first (fork pool), CUDA afterwards
    frames = make_frames(B, 1000 + rank * 100000, max(1, min(16, cores // max(world, 1))))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from geoflowslam_b200 import TrackingFrontend
    from geoflowslam_b200._lib import KP_DTYPE
    fe = TrackingFrontend(max_size=(W, H), max_batch=B, **ORB_CFG)
    S = fe.stride
    dev = torch.device("cuda", local)
    d_imgs = torch.from_numpy(frames).to(dev)
    d_out = dict(kp=torch.empty((B, S, 6), dtype=torch.float32, device=dev),
                 desc=torch.empty((B, S, 32), dtype=torch.uint8, device=dev),
                 n=torch.zeros(B, dtype=torch.int32, device=dev), mono=torch.zeros(B, dtype=torch.int32, device=dev),
                 train_idx=torch.empty((B, S), dtype=torch.int32, device=dev),
                 dist=torch.empty((B, S), dtype=torch.int32, device=dev),
                 inlier=torch.empty((B, S), dtype=torch.uint8, device=dev),
                 inlier_count=torch.zeros(B, dtype=torch.int32, device=dev))
    stream = torch.cuda.current_stream().cuda_stream

    def step_dev():
        fe.run_device(d_imgs, B, W, H, W, W * H, d_out, stream=stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        """k calls of fn between two CUDA events on the current stream -> ms (max over ranks)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        step_dev()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_dev, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * B / (ms_step / 1e3)

    # per-kernel durations (CUDA events recorded inside the C-ABI call, on the same stream)
    fe.set_profiling(True)
    prof = {}
    for _ in range(args.steps):
        step_dev()
        for k, v in fe.profile().items():
            prof.setdefault(k, []).append(v)
    fe.set_profiling(False)
    prof = {k: sum(v) / len(v) for k, v in prof.items()}
    n_host = d_out["n"].cpu().numpy()
    mean_kp = float(n_host.mean())

    # end to end through the host-buffer C-ABI call: pinned images in, all results out
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
    h_imgs_t = pin(frames.shape, torch.uint8)
    h_imgs_t.copy_(torch.from_numpy(frames))
    h_imgs = h_imgs_t.numpy()
    keep = []

    def pinned_np(shape, dt):
        nbytes = int(np.prod(shape)) * np.dtype(dt).itemsize
        t = torch.empty(max(nbytes, 1), dtype=torch.uint8).pin_memory()
        keep.append(t)
        return t.numpy()[:nbytes].view(dt).reshape(shape)

    h_out = fe.alloc_host_outputs(B, pinned_np)

    def step_e2e():
        fe.run(h_imgs, out=h_out, stream=stream)

    for _ in range(2):
        step_e2e()
    e2e_steps = max(1, min(args.steps, 5))
    ms_e2e = timed(step_e2e, e2e_steps) / e2e_steps
    e2e_value = world * B / (ms_e2e / 1e3)
    h2d = int(frames.nbytes)
    d2h = int(sum(v.nbytes for v in h_out.values()))
    assert np.array_equal(h_out["n"], n_host), "host-path and device-path results differ"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
    top = max(prof, key=prof.get)
    cand_per_frame = 3400.0  # measured mean FAST candidates/frame on this workload
    alg = algorithmic_bytes(top, B if top not in ("bf_hamming", "gms") else B - 1, mean_kp, cand_per_frame)
    achieved = alg / (prof[top] / 1e3) / 1e9
    # DRAM traffic per frame of the dominant kernels from the committed `ncu --set full` captures
    # (dram__bytes_read.sum + dram__bytes_write.sum; profiles/r01_summary.md), scaled to this launch
    ncu_traffic_per_frame = {"fast_cells": (113.192192e6 + 5.287680e6) / 128,     # profiles/r01_s3_k_fast_cells2_ncu_details.txt
                             "orient_desc": (122.779648e6 + 8.324864e6) / 128}   # profiles/r01_s3_k_orient_desc_ncu_details.txt
    ncu_traffic_src = {"fast_cells": "ncu capture profiles/r01_s3_k_fast_cells2_ncu_details.txt",
                       "orient_desc": "ncu capture profiles/r01_s3_k_orient_desc_ncu_details.txt"}
    traffic = ncu_traffic_per_frame[top] * B if top in ncu_traffic_per_frame else None
    # every stage against the same roof (algorithmic bytes / CUDA-event time), for the stage table in DESIGN.md
    stage_roof = {}
    for st_name, st_ms in prof.items():
        if st_ms <= 0.01:
            continue
        a = algorithmic_bytes(st_name, B if st_name not in ("bf_hamming", "gms") else B - 1, mean_kp, cand_per_frame)
        stage_roof[st_name] = {"ms": st_ms, "algorithmic_bytes": a, "achieved": a / (st_ms / 1e3) / 1e9, "frac": a / (st_ms / 1e3) / 1e9 / peak}
    roof = {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "stages": stage_roof,
            "frac": achieved / peak, "traffic": traffic, "traffic_source": ncu_traffic_src.get(top) if traffic else None,
            "algorithmic_bytes": alg, "ms_per_launch_group": prof[top], "peak_source": peak_src,
            "stage_ms": prof,
            "note": "every ORB stage is bound by issue slots (ncu: 63-85 % busy), not DRAM (3-13 % of peak); "
                    "DRAM traffic equals the algorithmic minimum (profiles/r01_summary.md)"}

    cpu = None
    if world == 1 and not args.no_cpu:
        # bounded sample: a short probe sizes it to about 8 s of wall on all host cores (capped by the batch)
        fps0, _ = cpu_frames_per_sec(frames[:min(B, max(2 * cores, 32))], cores)
        ns = int(min(B, max(2 * cores, 32, fps0 * 8.0)))
        fps, dt = cpu_frames_per_sec(frames[:ns], cores)
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": "first %d frames + %d pairs of the same batch, %.1f s wall, %d worker threads" % (ns, ns - 1, dt, cores)}

    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "configs[1]: ORB(1000 feats, 1.2, 8 levels, FAST 25/7) + BF-Hamming + GMS i->i+1",
                       "frames_per_gpu": B, "width": W, "height": H, "mean_keypoints": mean_kp,
                       "l2": "inputs larger than L2 (%.0f MB of frames per step)" % (frames.nbytes / 1e6),
                       "parallelism": "frames sharded across ranks, no data-path collective"},
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "frames/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": args.steps * fe.launches_per_call(B)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# Secondary workloads (BASELINE configs[2] GICP and configs[3] LocalInertialBA); same JSON contract
# ------------------------------------------------------------------------------------------------
def _clock_block(fn, local):
    import torch
    s = ClockSampler(local)
    s.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), s.stop()


def run_gicp(args):
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from geoflowslam_b200 import RegistrationGICP, synth
    from geoflowslam_b200.gicp import RESULT_DTYPE
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    P = args.batch if args.batch != 1024 else 512
    uniq = min(P, 32)  # distinct synthetic pairs; the batch tiles them (every pair is solved independently)
    with ThreadPoolExecutor(min(16, os.cpu_count() or 1)) as ex:
        pairs = list(ex.map(lambda i: synth.gicp_pair(2000 + i, n_target=50000), range(uniq)))
    stride = max(max(len(t), len(s)) for t, s, _ in pairs)
    tg = np.zeros((P, stride, 4), np.float32); sr = np.zeros((P, stride, 4), np.float32)
    nt = np.zeros(P, np.int32); ns = np.zeros(P, np.int32)
    for i in range(P):
        t, s_, _ = pairs[i % uniq]
        tg[i, :len(t)] = t; sr[i, :len(s_)] = s_; nt[i] = len(t); ns[i] = len(s_)
    T0 = np.tile(np.eye(4), (P, 1, 1))
    reg = RegistrationGICP(max_points=stride, max_pairs=P)
    dev = torch.device("cuda", local)
    d_tg, d_sr = torch.from_numpy(tg).to(dev), torch.from_numpy(sr).to(dev)
    d_nt, d_ns, d_T0 = torch.from_numpy(nt).to(dev), torch.from_numpy(ns).to(dev), torch.from_numpy(T0).to(dev)
    d_out = torch.zeros(P * RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    if args.gicp_track:
        # tracking mode: every step brings ONE new cloud per sequence (alternately the pair's source and target cloud) and
        # registers it against the previous step's, which is already preprocessed and resident
        calls = [0]
        reg.track_batch_device(d_tg, d_nt, P, stride, None, None, stream=stream)

        def step():
            calls[0] += 1
            new, n = (d_sr, d_ns) if calls[0] & 1 else (d_tg, d_nt)
            reg.track_batch_device(new, n, P, stride, d_T0, d_out, stream=stream)
    else:
        step = lambda: reg.align_batch_device(d_tg, d_nt, d_sr, d_ns, P, stride, d_T0, d_out, stream=stream)
    for _ in range(max(args.warmup, 1)):
        step()
    ms, clocks = _clock_block(lambda: [step() for _ in range(args.steps)], local)
    ms /= args.steps
    res = d_out.cpu().numpy().view(RESULT_DTYPE)
    M_t, M_s = float(res["n_target"].mean()), float(res["n_source"].mean())
    I, J = float(res["iterations"].mean() + 1), float(res["inner_evals"].mean())
    alg = P * (16 * (nt.mean() + ns.mean()) + 400 * (M_t + M_s) + M_s * (160 * I + 112 * J))  # SURVEY 8d A_gicp
    t0 = time.perf_counter(); res_h = reg.align_batch(tg, nt, sr, ns, T0); e2e_s = time.perf_counter() - t0
    if not args.gicp_track:
        assert np.array_equal(res_h["iterations"], res["iterations"])
    nsamp = min(uniq, 8)
    cpu_s = float("nan")
    if not args.no_cpu:
        from oracle import oracle as O
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max(1, (os.cpu_count() or 4) // 4)) as ex:  # 4 OpenMP threads per pair, as the reference
            list(ex.map(lambda i: O.gicp_align(pairs[i][0], pairs[i][1], threads=4), range(nsamp)))
        cpu_s = time.perf_counter() - t0
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    line = {"metric": "cloud pairs/sec GICP 50k-pt RGB-D pairs (BASELINE configs[2])", "value": P / (ms / 1e3), "unit": "pairs/s",
            "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[2]: RegistrationGICP voxel 0.02, max dist 0.1, k=10, <=20 LM iterations" +
                                   (" -- tracking mode: one new cloud per sequence and step, registered against the resident previous one" if args.gicp_track else ""),
                       "variant": {k: os.environ.get(k) for k in ("GFS_GICP_ORDER", "GFS_GICP_NN", "GFS_GICP_CELL", "GFS_GICP_KNN_CELLS") if os.environ.get(k)},
                       "pairs": P, "distinct_pairs": uniq, "points_per_cloud": int(nt.mean()), "downsampled": [M_t, M_s],
                       "outer_iterations": I, "inner_evals": J},
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "whole align call", "achieved": alg / (ms / 1e3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": alg / (ms / 1e3) / 1e9 / peak, "traffic": None},
            "cpu_baseline": {"value": nsamp / cpu_s, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": "%d pairs, 4 threads per pair" % nsamp},
            "e2e": {"value": P / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": int(tg.nbytes + sr.nbytes),
                    "d2h_bytes_per_step": int(res_h.nbytes)},
            "gpu_launches": args.steps * reg.last_launches()}
    print(json.dumps(line))


def run_ba(args, se3=False):
    import torch
    from geoflowslam_b200 import Optimizer, synth
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    B = args.batch if args.batch != 1024 else 256
    uniq = min(B, 8)
    if se3:  # Optimizer::LocalBundleAdjustment: 18 optimisable + 3 fixed keyframes, no inertial edges
        probs = [synth.lba_problem(seed=7000 + i, n_kf=18, n_fixed=3, n_points=3000) for i in range(uniq)]
    else:
        probs = [synth.ba_problem(seed=3000 + i) for i in range(uniq)]
    batch = [probs[i % uniq] for i in range(B)]
    opt = Optimizer(max_kf=21, max_points=3000, max_obs=16384, max_inertial=20, max_batch=B)
    stream = torch.cuda.current_stream().cuda_stream
    opt.upload(batch, stream)
    for _ in range(max(args.warmup, 1)):
        opt.solve_uploaded(stream)
    ms, clocks = _clock_block(lambda: [opt.solve_uploaded(stream) for _ in range(args.steps)], local)
    ms /= args.steps
    res = opt.download(stream)
    t0 = time.perf_counter(); opt.LocalInertialBA_batch(batch, stream); e2e_s = time.perf_counter() - t0
    one = Optimizer(max_kf=21, max_points=3000, max_obs=16384, max_inertial=20, max_batch=1)
    one.upload(batch[:1], stream); one.solve_uploaded(stream)
    ms1, _ = _clock_block(lambda: [one.solve_uploaded(stream) for _ in range(5)], local)
    from oracle import oracle as O
    t0 = time.perf_counter()
    for p in probs[:4]:
        O.ba_solve(p)
    cpu_s = (time.perf_counter() - t0) / 4
    trials = float(np.mean([r["lm_trials"] for r in res]))
    alg = B * trials * 7.6e6  # A_ba ~ 7.6 MB per LM trial
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    line = {"metric": ("problems/sec LocalBundleAdjustment 18+3 KF x 3000 MP (SURVEY 8f rank 3)" if se3 else
                       "problems/sec LocalInertialBA 20 KF x 3000 MP x 15k obs (BASELINE configs[3])"), "value": B / (ms / 1e3),
            "unit": "problems/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": ("LocalBundleAdjustment: 18 optimisable + 3 fixed keyframes, 3000 points, VertexSE3Expmap, 10 LM iterations" if se3 else
                                    "configs[3]: 20 KF, 3000 points, ~15k stereo/mono edges, 20 inertial edges, bLarge (4 its)"),
                       "batch": B, "distinct_problems": uniq, "lm_trials": trials, "single_problem_ms": ms1 / 5},
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "whole solve", "achieved": alg / (ms / 1e3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": alg / (ms / 1e3) / 1e9 / peak, "traffic": None},
            "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "problems/s", "cores": 1, "kind": "port", "sample": "4 problems, 1 thread (g2o OpenMP is off)"},
            "e2e": {"value": B / e2e_s, "unit": "problems/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None},
            "gpu_launches": args.steps * opt.last_launches()}
    print(json.dumps(line))


def run_pose(args):
    """Optimizer::PoseOptimization, one problem per frame, batched."""
    import torch
    from geoflowslam_b200 import PoseOptimizer, synth
    from geoflowslam_b200 import pose as pose_mod
    import ctypes as C
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    B = args.batch
    uniq = min(B, 64)
    probs = [synth.pose_problem(seed=5000 + i, n_obs=400) for i in range(uniq)]
    batch = [probs[i % uniq] for i in range(B)]
    opt = PoseOptimizer(max_obs=512, max_batch=B)
    # pack once: the timed call is the C-ABI entry point with host pointers (H2D + kernel + D2H)
    Ps = (pose_mod.PoseProblem * B)()
    Rs = (pose_mod.PoseResult * B)()
    keep = [pose_mod.pack_problem(pr, Ps[i])[1] for i, pr in enumerate(batch)]
    outs = [pose_mod.alloc_result(400, Rs[i])[1] for i in range(B)]
    stream = torch.cuda.current_stream().cuda_stream
    call = lambda: pose_mod.check(opt._L.gfs_pose_optimize_batch(opt._h, stream, Ps, B, Rs))
    for _ in range(max(args.warmup, 3)):
        call()
    ms, clocks = _clock_block(lambda: [call() for _ in range(args.steps)], local)
    ms /= args.steps
    from oracle import oracle as O
    t0 = time.perf_counter()
    for p in probs[:32]:
        O.pose_optimize(p)
    cpu_s = (time.perf_counter() - t0) / 32
    iters = float(np.mean([sum(Rs[i].lm_iterations) for i in range(B)]))
    h2d = B * (56 + 400 * (24 + 12 + 4))
    d2h = B * (96 + 400 * 5)
    alg = B * iters * 2.5 * 400 * (24 + 12 + 4 + 24)  # per LM iteration ~2.5 passes over (Xw, uvr, info, stored error)
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    line = {"metric": "frames/sec PoseOptimization 400 map-point observations per frame (SURVEY 8f rank 1)",
            "value": B / (ms / 1e3), "unit": "frames/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "PoseOptimization: 4 rounds x <=10 LM iterations, 400 observations (80 % stereo), 10 % gross outliers",
                       "batch": B, "distinct_problems": uniq, "mean_lm_iterations": iters,
                       "note": "value is measured through the host-pointer C-ABI call (the only entry point): it equals e2e"},
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_pose_opt", "achieved": alg / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms / 1e3) / 1e9 / peak, "traffic": None},
            "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "frames/s", "cores": 1, "kind": "port",
                             "sample": "32 frames, 1 thread (the tracking thread runs it serially)"},
            "e2e": {"value": B / (ms / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": args.steps * opt.last_launches()}
    print(json.dumps(line))


def run_pose_inertial(args):
    """Optimizer::PoseInertialOptimizationLastFrame / LastKeyFrame, one problem per frame, batched."""
    import torch
    from geoflowslam_b200 import PoseInertialOptimizer, synth
    from geoflowslam_b200 import pose_inertial as pin
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    B = args.batch
    uniq = min(B, 64)
    # TrackLocalMap alternates: LastFrame while the map is unchanged, LastKeyFrame after a map update (Tracking.cc:3770-3795)
    probs = [synth.pose_inertial_problem(seed=6000 + i, mode=(0 if i % 4 == 0 else 1), n_obs=400) for i in range(uniq)]
    batch = [probs[i % uniq] for i in range(B)]
    opt = PoseInertialOptimizer(max_obs=512, max_batch=B)
    Ps = (pin.PoseInertialProblem * B)()
    Rs = (pin.PoseInertialResult * B)()
    keep = [pin.pack_problem(pr, Ps[i])[1] for i, pr in enumerate(batch)]
    outs = [pin.alloc_result(400, Rs[i])[1] for i in range(B)]
    stream = torch.cuda.current_stream().cuda_stream
    call = lambda: pin.check(opt._L.gfs_pose_inertial_optimize_batch(opt._h, stream, Ps, B, Rs))
    for _ in range(max(args.warmup, 3)):
        call()
    ms, clocks = _clock_block(lambda: [call() for _ in range(args.steps)], local)
    ms /= args.steps
    from oracle import oracle as O
    t0 = time.perf_counter()
    for p in probs[:32]:
        O.pose_inertial_optimize(p)
    cpu_s = (time.perf_counter() - t0) / 32
    iters = float(np.mean([sum(Rs[i].gn_iterations) for i in range(B)]))
    h2d = B * (5400 + 400 * (24 + 12 + 4 + 1))
    d2h = B * (2000 + 400 * 5)
    alg = B * iters * 400 * (24 + 12 + 4 + 24)  # per GN iteration one pass over (Xw, uvr, info, stored error)
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    line = {"metric": "frames/sec PoseInertialOptimizationLast{Frame,KeyFrame} 400 map-point observations per frame (SURVEY 8f rank 1)",
            "value": B / (ms / 1e3), "unit": "frames/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "PoseInertialOptimization: 4 rounds x 10 Gauss-Newton iterations, 15 (LastKeyFrame) / 30 (LastFrame) unknowns, 400 observations (80 % stereo), 10 % gross outliers",
                       "batch": B, "distinct_problems": uniq, "mean_gn_iterations": iters,
                       "note": "latency-bound dense solves, not an HBM kernel; value is measured through the host-pointer C-ABI call (the only entry point): it equals e2e"},
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_pose_inertial", "achieved": alg / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms / 1e3) / 1e9 / peak, "traffic": None},
            "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "frames/s", "cores": 1, "kind": "port",
                             "sample": "32 frames, 1 thread (the tracking thread runs it serially)"},
            "e2e": {"value": B / (ms / 1e3), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": args.steps * opt.last_launches()}
    print(json.dumps(line))


def run_klt(args):
    """cv::buildOpticalFlowPyramid per frame + ORBmatcher::fbKltTracking per consecutive pair."""
    import cv2
    import torch
    from geoflowslam_b200 import KltTracker
    local = int(os.environ.get("LOCAL_RANK", "0"))
    B = args.batch
    cores = os.cpu_count() or 1
    uniq = min(B, 64)
    frames_u = make_frames(uniq, 1000, min(16, cores))
    NP = 1000
    kps_u = np.zeros((uniq, 1024, 2), np.float32); n_u = np.zeros(uniq, np.int32)
    for i in range(uniq):
        p = cv2.goodFeaturesToTrack(frames_u[i], NP, 0.01, 5).reshape(-1, 2).astype(np.float32)
        kps_u[i, :len(p)] = p; n_u[i] = len(p)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    idx = np.arange(B) % uniq
    # frame i is tracked into frame i+1 of the same 8-frame scene group (the last frame of a group into its first)
    nxt = (idx // 8) * 8 + (idx % 8 + 1) % 8
    frames = np.ascontiguousarray(frames_u[idx])
    d_imgs = torch.from_numpy(frames).to(dev)
    d_kps = torch.from_numpy(np.ascontiguousarray(kps_u[idx])).to(dev)
    d_n = torch.from_numpy(np.ascontiguousarray(n_u[idx])).to(dev)
    d_next_idx = torch.from_numpy(nxt.astype(np.int64)).to(dev)
    trk = KltTracker(max_size=(W, H), levels=3, max_points=1024, max_batch=B)
    pb = trk.pyramid_bytes(W, H)
    d_pyr = torch.zeros((B, pb), dtype=torch.uint8, device=dev)
    d_cur = torch.zeros((B, pb), dtype=torch.uint8, device=dev)
    d_pr = d_kps.clone(); d_st = torch.zeros((B, 1024), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_pyr, t_trk = [], []

    def step(timed=False):
        d_pr.copy_(d_kps)
        if timed: ev[0].record()
        trk.build_pyramids_device(d_imgs, B, W, H, W, W * H, d_pyr, stream=stream)
        torch.index_select(d_pyr, 0, d_next_idx, out=d_cur)   # pairing only: the "current" frame of pair i
        if timed: ev[1].record()
        trk.fb_track_device(d_pyr, d_cur, B, W, H, d_kps, d_pr, d_n, 1024, d_st, nwinsize=35, nbpyrlvl=3, ferr=15.0, fmax_fbklt_dist=0.5,
                            stream=stream)
        if timed:
            ev[2].record(); torch.cuda.synchronize()
            t_trk.append(ev[1].elapsed_time(ev[2]))

    for _ in range(max(args.warmup, 3)):
        step()
    ms, clocks = _clock_block(lambda: [step() for _ in range(args.steps)], local)
    ms /= args.steps
    for _ in range(3):
        step(True)
    trk_ms = sum(t_trk) / len(t_trk)
    tracked = float(d_st.sum().item()) / float(n_u[idx].sum())
    # e2e: host images + keypoints in, tracks out, through the host-pointer call (one pair per call)
    one = KltTracker(max_size=(W, H), levels=3, max_points=1024, max_batch=1)
    ne = min(B, 64)
    one.fbKltTracking(frames[0], frames_u[nxt[0]], kps_u[idx[0], :n_u[idx[0]]], kps_u[idx[0], :n_u[idx[0]]])
    t0 = time.perf_counter()
    for i in range(ne):
        k = kps_u[idx[i], :n_u[idx[i]]]
        one.fbKltTracking(frames[i], frames_u[nxt[i]], k, k)
    e2e_s = (time.perf_counter() - t0) / ne
    # CPU arm: the OpenCV calls the reference makes (cv2 wheel, all host threads)
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
    fl = cv2.OPTFLOW_USE_INITIAL_FLOW + cv2.OPTFLOW_LK_GET_MIN_EIGENVALS

    def cpu_pair(i):
        a, b = frames_u[i], frames_u[(i // 8) * 8 + (i % 8 + 1) % 8]
        k = kps_u[i, :n_u[i]]
        _, pa = cv2.buildOpticalFlowPyramid(a, (35, 35), 3)
        _, pbb = cv2.buildOpticalFlowPyramid(b, (35, 35), 3)
        pr, st, er = cv2.calcOpticalFlowPyrLK(a, b, k, k.copy(), winSize=(35, 35), maxLevel=3, criteria=crit, flags=fl)
        g = st.ravel().astype(bool) & ~(er.ravel() > 15.0)
        if g.any():
            cv2.calcOpticalFlowPyrLK(b, a, pr[g], k[g].copy(), winSize=(35, 35), maxLevel=0, criteria=crit, flags=fl)

    cpu_pair(0)
    t0 = time.perf_counter()
    cnt = 0
    while time.perf_counter() - t0 < 8.0:
        cpu_pair(cnt % uniq); cnt += 1
    cpu_s = (time.perf_counter() - t0) / cnt
    npts = float(n_u[idx].mean())
    alg = B * (W * H * 1.33 * (1 + 1 + 4)) + B * npts * 4 * (37 * 37 * 5 + 6 * 36 * 36)  # pyramid r/w + per point and level: template + ~6 windows
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    line = {"metric": "frame pairs/sec buildOpticalFlowPyramid + fbKltTracking, VGA, ~%d points per pair (SURVEY 8f rank 2)" % int(npts),
            "value": B / (ms / 1e3), "unit": "pairs/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i16/f32", "data": "synthetic",
            "config": {"workload": "optical-flow front end: 3-level pyramid + Scharr derivatives per frame, forward LK (35x35, 4 levels, <=30 its) + backward LK (level 0) per point",
                       "pairs": B, "distinct_frames": uniq, "points_per_pair": npts, "tracked_fraction": tracked,
                       "stage_ms": {"track": trk_ms, "pyramid_and_pairing": ms - trk_ms}},
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_klt_track", "achieved": alg / (ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms / 1e3) / 1e9 / peak, "traffic": None,
                         "note": "instruction-bound fixed-point bilinear sampling out of shared memory / L1, not an HBM kernel"},
            "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "pairs/s", "cores": cores, "kind": "reference",
                             "sample": "%d pairs in %.1f s: cv2 %s buildOpticalFlowPyramid + calcOpticalFlowPyrLK forward/backward, the OpenCV calls the reference makes (OpenCV's own thread pool)" % (cnt, cnt * cpu_s, cv2.__version__)},
            "e2e": {"value": 1.0 / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": 2 * W * H + int(npts) * 16, "d2h_bytes_per_step": int(npts) * 9,
                    "note": "one pair per host-pointer call (gfs_klt_fb_track), both pyramids rebuilt per call"},
            "gpu_launches": args.steps * (5 + 1 + 1)}
    print(json.dumps(line))


def run_track(args):
    """BASELINE.json's metric shape -- "frames/sec VGA RGBD-I (1k feats) track+LocalBA" -- as an OPEN-LOOP composition of the
    built stages on synthetic inputs of the named shapes: one step advances B independent sequences by one frame each
    (ORB extraction + BF/GMS, optical-flow pyramid + fbKltTracking, PoseInertialOptimizationLastFrame, depth -> cloud +
    RegistrationGICP on 50k-point clouds) and runs LocalInertialBA for the B/10 sequences that insert a keyframe.  The
    stages are not causally chained (the Tracking / LocalMapping state machines are out of scope); every
    stage is the parity-tested kernel path of its own workload."""
    import cv2
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from geoflowslam_b200 import KltTracker, Optimizer, PoseInertialOptimizer, RegistrationGICP, TrackingFrontend, synth
    from geoflowslam_b200 import pose_inertial as pin
    from geoflowslam_b200._lib import check, lib, ptr
    from geoflowslam_b200.gicp import RESULT_DTYPE
    local = int(os.environ.get("LOCAL_RANK", "0"))
    B = args.batch if args.batch != 1024 else 128
    cores = os.cpu_count() or 1
    uniq = min(B, 32)
    frames_u = make_frames(uniq, 1000, min(16, cores))
    with ThreadPoolExecutor(min(16, cores)) as ex:
        pairs = list(ex.map(lambda i: synth.gicp_pair(2000 + i, n_target=50000), range(8)))
    ba_probs = [synth.ba_problem(seed=3000 + i) for i in range(2)]
    pin_probs = [synth.pose_inertial_problem(seed=6000 + i, mode=1, n_obs=400) for i in range(16)]
    kps_u = np.zeros((uniq, 1024, 2), np.float32); n_u = np.zeros(uniq, np.int32)
    for i in range(uniq):
        p = cv2.goodFeaturesToTrack(frames_u[i], 1000, 0.01, 5).reshape(-1, 2).astype(np.float32)
        kps_u[i, :len(p)] = p; n_u[i] = len(p)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream().cuda_stream
    idx = np.arange(B) % uniq
    frames = np.ascontiguousarray(frames_u[idx])
    d_imgs = torch.from_numpy(frames).to(dev)
    # -- ORB + match
    fe = TrackingFrontend(max_size=(W, H), max_batch=B, **ORB_CFG)
    S = fe.stride
    d_out = dict(kp=torch.empty((B, S, 6), dtype=torch.float32, device=dev), desc=torch.empty((B, S, 32), dtype=torch.uint8, device=dev),
                 n=torch.zeros(B, dtype=torch.int32, device=dev), mono=torch.zeros(B, dtype=torch.int32, device=dev),
                 train_idx=torch.empty((B, S), dtype=torch.int32, device=dev), dist=torch.empty((B, S), dtype=torch.int32, device=dev),
                 inlier=torch.empty((B, S), dtype=torch.uint8, device=dev), inlier_count=torch.zeros(B, dtype=torch.int32, device=dev))
    # -- optical flow
    trk = KltTracker(max_size=(W, H), levels=3, max_points=1024, max_batch=B)
    pb = trk.pyramid_bytes(W, H)
    d_pyr = torch.zeros((B, pb), dtype=torch.uint8, device=dev); d_cur = torch.zeros((B, pb), dtype=torch.uint8, device=dev)
    nxt = torch.from_numpy(((idx // 8) * 8 + (idx % 8 + 1) % 8).astype(np.int64)).to(dev)
    d_kps = torch.from_numpy(np.ascontiguousarray(kps_u[idx])).to(dev); d_pr = d_kps.clone()
    d_nk = torch.from_numpy(np.ascontiguousarray(n_u[idx])).to(dev); d_st = torch.zeros((B, 1024), dtype=torch.uint8, device=dev)
    # -- pose-inertial (host-pointer C ABI: packed once)
    pio = PoseInertialOptimizer(max_obs=512, max_batch=B)
    Ps = (pin.PoseInertialProblem * B)(); Rs = (pin.PoseInertialResult * B)()
    keep = [pin.pack_problem(pin_probs[i % len(pin_probs)], Ps[i])[1] for i in range(B)]
    outs = [pin.alloc_result(400, Rs[i])[1] for i in range(B)]
    # -- depth -> cloud + GICP
    depth = torch.rand((B, H, W), dtype=torch.float32, device=dev) * 5.0 + 0.5
    d_cloud = torch.empty((B, 65536, 4), dtype=torch.float32, device=dev); d_cn = torch.zeros(B, dtype=torch.int32, device=dev)
    stride = max(max(len(t), len(s_)) for t, s_, _ in pairs)
    tg = np.zeros((B, stride, 4), np.float32); sr = np.zeros((B, stride, 4), np.float32); nt = np.zeros(B, np.int32); ns = np.zeros(B, np.int32)
    for i in range(B):
        t, s_, _ = pairs[i % len(pairs)]
        tg[i, :len(t)] = t; sr[i, :len(s_)] = s_; nt[i] = len(t); ns[i] = len(s_)
    reg = RegistrationGICP(max_points=stride, max_pairs=B)
    d_tg, d_sr = torch.from_numpy(tg).to(dev), torch.from_numpy(sr).to(dev)
    d_nt, d_ns = torch.from_numpy(nt).to(dev), torch.from_numpy(ns).to(dev)
    d_T0 = torch.from_numpy(np.tile(np.eye(4), (B, 1, 1))).to(dev)
    d_res = torch.zeros(B * RESULT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    # -- LocalInertialBA for the sequences that insert a keyframe (1 in 10)
    nba = max(1, B // 10)
    opt = Optimizer(max_kf=21, max_points=3000, max_obs=16384, max_inertial=20, max_batch=nba)
    opt.upload([ba_probs[i % 2] for i in range(nba)], stream)
    L = lib()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
    stage = {}

    def step(timed=False):
        def mark(i):
            if timed: ev[i].record()
        mark(0)
        fe.run_device(d_imgs, B, W, H, W, W * H, d_out, stream=stream)
        mark(1)
        trk.build_pyramids_device(d_imgs, B, W, H, W, W * H, d_pyr, stream=stream)
        torch.index_select(d_pyr, 0, nxt, out=d_cur)
        d_pr.copy_(d_kps)
        trk.fb_track_device(d_pyr, d_cur, B, W, H, d_kps, d_pr, d_nk, 1024, d_st, stream=stream)
        mark(2)
        pin.check(pio._L.gfs_pose_inertial_optimize_batch(pio._h, stream, Ps, B, Rs))
        mark(3)
        check(L.gfs_depth_to_cloud_batch_device(stream, ptr(depth), B, W, H, W, W * H, 2, 606.986, 607.011, 311.519, 247.260, ptr(d_cloud), 65536, ptr(d_cn)))
        reg.align_batch_device(d_tg, d_nt, d_sr, d_ns, B, stride, d_T0, d_res, stream=stream)
        mark(4)
        opt.solve_uploaded(stream)
        mark(5)
        if timed:
            torch.cuda.synchronize()
            for k, i in (("orb_match", 0), ("klt", 1), ("pose_inertial", 2), ("depth_cloud_gicp", 3), ("local_inertial_ba", 4)):
                stage.setdefault(k, []).append(ev[i].elapsed_time(ev[i + 1]))

    # The sequences are independent and so are the stages of this open-loop composition: the tracking front end (ORB +
    # match, optical flow) stays on the main stream, the pose optimiser, depth -> cloud + GICP and the local-mapping BA
    # get a stream and a host thread each (their C-ABI calls block on their own stream; ctypes releases the GIL) -- the
    # arrangement of the reference's Tracking / LocalMapping threads.  GICP's divergence-bound kernels leave issue slots
    # the other stages' kernels can use.  GFS_TRACK_SEQUENTIAL=1 times the one-stream step instead.
    s_main = torch.cuda.current_stream()
    side = [torch.cuda.Stream(device=dev) for _ in range(3)]
    pool = ThreadPoolExecutor(3)

    def _gicp(cs):
        torch.cuda.set_device(local)
        check(L.gfs_depth_to_cloud_batch_device(cs, ptr(depth), B, W, H, W, W * H, 2, 606.986, 607.011, 311.519, 247.260, ptr(d_cloud), 65536, ptr(d_cn)))
        reg.align_batch_device(d_tg, d_nt, d_sr, d_ns, B, stride, d_T0, d_res, stream=cs)

    def _ba(cs):
        torch.cuda.set_device(local)
        opt.solve_uploaded(cs)

    def _pose(cs):
        torch.cuda.set_device(local)
        pin.check(pio._L.gfs_pose_inertial_optimize_batch(pio._h, cs, Ps, B, Rs))

    def step_concurrent():
        for cs in side:
            cs.wait_stream(s_main)
        futs = [pool.submit(_gicp, side[0].cuda_stream), pool.submit(_ba, side[1].cuda_stream), pool.submit(_pose, side[2].cuda_stream)]
        fe.run_device(d_imgs, B, W, H, W, W * H, d_out, stream=stream)
        trk.build_pyramids_device(d_imgs, B, W, H, W, W * H, d_pyr, stream=stream)
        torch.index_select(d_pyr, 0, nxt, out=d_cur)
        d_pr.copy_(d_kps)
        trk.fb_track_device(d_pyr, d_cur, B, W, H, d_kps, d_pr, d_nk, 1024, d_st, stream=stream)
        for f in futs:
            f.result()
        for cs in side:
            s_main.wait_stream(cs)

    for _ in range(max(args.warmup, 3)):
        step()
    ms_seq, clocks = _clock_block(lambda: [step() for _ in range(args.steps)], local)
    ms_seq /= args.steps
    ms, mode = ms_seq, "one stream, stages back to back"
    if not os.environ.get("GFS_TRACK_SEQUENTIAL"):
        try:
            for _ in range(2):
                step_concurrent()
            ms_c, clocks_c = _clock_block(lambda: [step_concurrent() for _ in range(args.steps)], local)
            ms_c /= args.steps
            if ms_c < ms_seq:
                ms, clocks, mode = ms_c, clocks_c, "front end on the main stream; pose optimiser, GICP and BA on a stream + host thread each"
        except Exception as e:  # keep the verified one-stream number rather than no number
            mode = "one stream, stages back to back (concurrent step failed: %s)" % e
    pool.shutdown()
    for _ in range(2):
        step(True)
    stage = {k: sum(v) / len(v) for k, v in stage.items()}
    # ---- CPU arm: core-seconds per frame of the restated reference path, stage by stage, one thread each (bounded samples);
    # reported as cores / core-seconds, i.e. assuming the CPU scales perfectly over independent sequences
    from oracle import oracle as O
    cv2.setNumThreads(1)
    t0 = time.perf_counter(); cpu_frames_per_sec(frames_u[:8], 1); c_orb = (time.perf_counter() - t0) / 8
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01); fl = cv2.OPTFLOW_USE_INITIAL_FLOW + cv2.OPTFLOW_LK_GET_MIN_EIGENVALS
    t0 = time.perf_counter()
    for i in range(8):
        a, b = frames_u[i], frames_u[(i // 8) * 8 + (i % 8 + 1) % 8]; k = kps_u[i, :n_u[i]]
        cv2.buildOpticalFlowPyramid(b, (35, 35), 3)
        pr, st, er = cv2.calcOpticalFlowPyrLK(a, b, k, k.copy(), winSize=(35, 35), maxLevel=3, criteria=crit, flags=fl)
        g = st.ravel().astype(bool)
        cv2.calcOpticalFlowPyrLK(b, a, pr[g], k[g].copy(), winSize=(35, 35), maxLevel=0, criteria=crit, flags=fl)
    c_klt = (time.perf_counter() - t0) / 8
    cv2.setNumThreads(-1)
    t0 = time.perf_counter()
    for p in pin_probs[:8]:
        O.pose_inertial_optimize(p)
    c_pin = (time.perf_counter() - t0) / 8
    t0 = time.perf_counter(); O.gicp_align(pairs[0][0], pairs[0][1], threads=1); O.gicp_align(pairs[1][0], pairs[1][1], threads=1)
    c_gicp = (time.perf_counter() - t0) / 2
    t0 = time.perf_counter(); O.ba_solve(ba_probs[0]); c_ba = (time.perf_counter() - t0) / 10
    core_s = dict(orb_match=c_orb, klt=c_klt, pose_inertial=c_pin, gicp=c_gicp, local_inertial_ba_per_frame=c_ba)
    cpu_fps = cores / sum(core_s.values())
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0)) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    res = d_res.cpu().numpy().view(RESULT_DTYPE)
    M_t, M_s = float(res["n_target"].mean()), float(res["n_source"].mean())
    I, J = float(res["iterations"].mean() + 1), float(res["inner_evals"].mean())
    alg = B * (16 * (nt.mean() + ns.mean()) + 400 * (M_t + M_s) + M_s * (160 * I + 112 * J))
    line = {"metric": "frames/sec VGA RGBD-I (1k feats) track+LocalBA, open-loop composition of the built stages (BASELINE metric shape)",
            "value": B / (ms / 1e3), "unit": "frames/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/f32/f64", "data": "synthetic",
            "config": {"workload": "per frame: ORB(1000)+BF/GMS, KLT pyramid + fbKltTracking, PoseInertialOptimizationLastFrame (400 obs), depth->cloud + GICP (50k-pt pair); per 10 frames: LocalInertialBA (20 KF x 3000 MP)",
                       "sequences": B, "keyframe_every": 10, "stage_ms": stage, "ms_per_step_one_stream": ms_seq, "step_mode": mode,
                       "note": "stages are not causally chained (open loop); stage_ms are one-stream CUDA-event times"},
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "GICP align (dominant stage)", "achieved": alg / (stage["depth_cloud_gicp"] / 1e3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": alg / (stage["depth_cloud_gicp"] / 1e3) / 1e9 / peak, "traffic": None},
            "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": "core-seconds per frame, one thread per stage on bounded samples: %s; value = cores / sum (perfect scaling assumed)" % json.dumps({k: round(v, 5) for k, v in core_s.items()})},
            "e2e": {"value": None, "unit": "frames/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                    "note": "inputs are device-resident except the pose-inertial problems (host pointers); see the per-workload e2e numbers"},
            "gpu_launches": None}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--gicp-track", action="store_true", help="--workload gicp in tracking mode (gfs_gicp_track_batch_device)")
    ap.add_argument("--workload", default="orb", choices=["orb", "gicp", "ba", "lba", "pose", "pose_inertial", "klt", "track"],
                    help="orb = BASELINE configs[1] (the driver's default); gicp / ba = configs[2] / configs[3]")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "gicp":
        run_gicp(args)
    elif args.workload == "ba":
        run_ba(args)
    elif args.workload == "lba":
        run_ba(args, se3=True)
    elif args.workload == "pose":
        run_pose(args)
    elif args.workload == "pose_inertial":
        run_pose_inertial(args)
    elif args.workload == "klt":
        run_klt(args)
    elif args.workload == "track":
        run_track(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
