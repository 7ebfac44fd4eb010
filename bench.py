#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native EZCalib hot path.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line.  A "step" is one Levenberg-Marquardt iteration (Schur pieces for the current radius, linear solve of the reduced
system, back-substitution, speculative evaluation of cost + Jacobian + normal equations at the candidate,
accept/reject) over the cfg5 workload of BASELINE.json: 10,000,080 observations per GPU (69,445 views x 144 AprilGrid
corners), for each of the 7 distortion models (mono focal).  Exactly K iterations are timed per model for any K (a
long solve that converges is continued from the initial parameters, see solve_exactly).

  value   = 10M-observation LM iterations per second, whole job, inputs resident in HBM
            (observations processed per second / 10,000,080; aggregate over the 7 models)
  e2e     = the same through the host-buffer C ABI (ezc_lm_create + set_params + solve + get_params):
            pinned host buffers, H2D of observations/poses and D2H of the results inside the timed region
            (median of three identical solves: one solve is ~12 ms of wall time)
  roofline= the dominant kernel (lm_accumulate) against the FP64 DFMA peak measured in the same run
  cpu_baseline = the CPU oracle ("restated Ceres", oracle/lm_oracle.cpp) on a bounded sample
  detect / detect_chessboard / detect_chessboard_stereo = images/s detect+subpix for BASELINE.json's cfg3 (AprilGrid
            2448x2048), cfg1 (chessboard 1280x1024, rad-tan 5) and cfg2 (stereo chessboard 1920x1080, Kannala-Brandt)
            geometries, each with its own e2e (pinned host images, asynchronous upload), HBM roofline and the reference
            (the reference's own AprilTag code / cv2 with the reference's flags) timed on the host cores with a parity
            check on the sample

`--impl reference` times the reference's CPU path: Ceres itself is not available in this image
(SURVEY.md 8c), so this is the oracle port with all host threads on a bounded sample.
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

N_VIEWS_10M = 69445
N_PTS = 144
OBS_10M = N_VIEWS_10M * N_PTS  # 10,000,080
IMG_W, IMG_H = 2448, 2048
BENCH_MODELS = ("rad1", "rad2", "rad3", "radtan4", "radtan5", "radtan8", "kannalabrandt")

# Algorithmic FLOPs per observation of one LM evaluation (DESIGN.md "LM roofline"): residual +
# analytic Jacobian + Cauchy weight (F_rj, counted from lm_models.cuh) plus the Gram accumulation
# 4*(d(d+1)/2 + d) + 4 with d = 6 + k, k = nf + 2 + nd.
F_RJ = {"rad1": 165, "rad2": 171, "rad3": 179, "radtan4": 203, "radtan5": 211, "radtan8": 262, "kannalabrandt": 231}


def flops_per_obs(model: str, mono: bool = True) -> int:
    from ezcalib_b200 import synth
    nd = synth.NUM_DIST[synth.MODEL_ID[model]]
    k = (1 if mono else 2) + 2 + nd
    d = 6 + k
    return F_RJ[model] + 4 * (d * (d + 1) // 2 + d) + 4


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_problem(model: str, n_views: int, seed_offset: int = 0):
    from ezcalib_b200 import synth
    tgt = synth.aprilgrid_target(6, 6, 0.06, 0.3)
    return synth.make_lm_problem(model, True, n_views, tgt, IMG_W, IMG_H, seed=synth.MASTER_SEED + seed_offset,
                                 roll_range=np.pi)


def measured_peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            pass
    return {}


def measured_traffic() -> dict:
    """DRAM traffic measured with ncu (dram__bytes_read.sum + dram__bytes_write.sum), written by tools/ncu_traffic.py into
    profiles/ncu_traffic_r02.json -- measurements, not literals; absent file => traffic is reported as null."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            pass
    return {}


# ------------------------------------------------------------------------------------------- CPU arm
def run_cpu(args, models, n_views_sample: int, as_reference: bool):
    from ezcalib_b200 import synth
    from oracle import lm_oracle as O
    cores = host_cores()
    tot_iters, tot_time = 0, 0.0
    detail = {}
    for mi, model in enumerate(models):
        P = make_problem(model, n_views_sample, seed_offset=mi)
        mid = synth.MODEL_ID[model]
        # warm-up: one short solve
        O.solve(mid, True, P.obs, P.target, P.width, P.height, P.focal_init, P.pp_init, P.dist_init, P.poses_init,
                max_iterations=max(1, args.warmup), threads=cores, trace=False)
        t0 = time.perf_counter()
        res = O.solve(mid, True, P.obs, P.target, P.width, P.height, P.focal_init, P.pp_init, P.dist_init, P.poses_init,
                      max_iterations=args.steps, threads=cores, trace=False)
        dt = time.perf_counter() - t0
        iters = res[4]["iterations"] - 1
        tot_iters += iters
        tot_time += dt
        detail[model] = round(iters / dt * (n_views_sample * N_PTS) / OBS_10M, 6)
    value = tot_iters / tot_time * (n_views_sample * N_PTS) / OBS_10M
    sample = f"{n_views_sample} views x {N_PTS} pts = {n_views_sample * N_PTS} obs per model, {len(models)} models, {args.steps} LM iterations each (oracle stops early if it converges)"
    return value, tot_time / max(tot_iters, 1) * 1e3, dict(value=value, unit="LM iters/s (10M obs)", cores=cores, kind="port",
                                                           sample=sample, per_model=detail)


# ------------------------------------------------------------------------------------------- GPU arm
def solve_exactly(solver, P, K, fixed, poses=None):
    """Exactly K LM iterations of the workload.  A long run converges: steps start to be rejected, the trust region
    shrinks to its minimum and the solve ends (termination 4) before K.  Then the remaining iterations continue from the
    initial parameters again -- every timed step stays a full iteration on the full problem; the restart (one upload of
    the poses and one more initial evaluation) is paid inside the timed region."""
    done, jac, acc, last = 0, 0, 0, None
    while done < K:
        summ, _ = solver.solve(max_num_iterations=K - done, **fixed)
        it = summ["iterations"] - 1
        if it < 1:
            raise RuntimeError(f"LM solve made no iteration: {summ}")
        done += it
        jac += summ["jacobian_evaluations"]
        acc += summ["num_successful"] - 1
        last = summ
        if done < K:
            solver.set_params(P.focal_init, P.pp_init, P.dist_init, P.poses_init if poses is None else poses)
    out = dict(last)
    out["iterations"], out["jacobian_evaluations"], out["num_successful"] = done + 1, jac, acc + 1
    return out


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from ezcalib_b200 import lm, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.current_stream()
    ctx = lm.Context(local_rank, stream=stream.cuda_stream)

    from ezcalib_b200 import dist as ezdist
    exchange = "none"
    if world > 1:
        if os.environ.get("EZC_BENCH_NCCL"):
            # round-1 path: the packets go through torch.distributed's NCCL all-reduce via a Python callback
            ctx.set_allreduce(ezdist.make_allreduce(f"cuda:{local_rank}"), rank, world)
            exchange = "torch.distributed all_reduce (NCCL) through the ezc_set_allreduce callback"
        else:
            # the library's own exchange: mailboxes in peer memory, IPC handles all-gathered once through torch.distributed
            ezdist.init_peer_exchange(ctx, rank, world)
            exchange = "ezc_comm: one-shot peer-memory all-reduce over NVLink, fixed rank order (csrc/comm.cu)"

    if args.only_rig:
        rig = run_cfg4(ctx, args, rank, world, stream)
        if rank == 0:
            emit({"metric": "cfg4 only (development run)", "cfg4": rig})
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        ctx.close()
        return
    models = BENCH_MODELS if not args.models else tuple(args.models.split(","))
    n_views = args.views
    K, W = args.steps, args.warmup
    fixed = dict(function_tolerance=-1.0, parameter_tolerance=-1.0, gradient_tolerance=-1.0)  # never terminate early: exactly K iterations

    # ---- per-model device-resident timing (inputs already in HBM) and e2e (host buffers)
    peak_fp64 = ctx.measure_fp64_peak()
    peak_dmma = ctx.measure_dmma_peak()
    sampler = ClockSampler(local_rank)
    dev_ms, e2e_ms, jac_evals, detail = 0.0, 0.0, 0, {}
    h2d_bytes = d2h_bytes = 0
    launches0 = None
    acc_ms_total, acc_flops_total = 0.0, 0.0
    sampler.start()
    for mi, model in enumerate(models):
        P = make_problem(model, n_views, seed_offset=mi + 100 * rank)
        mid = synth.MODEL_ID[model]
        # pinned host copies (what a caller of the C ABI holds)
        obs_pin = torch.from_numpy(P.obs).pin_memory()
        poses_pin = torch.from_numpy(P.poses_init).pin_memory()
        # -------- device-resident: observations uploaded once, outside the timed region
        obs_dev = obs_pin.cuda(non_blocking=True)
        pts_dev = torch.from_numpy(P.target).cuda()
        solver = lm.LmSolver(ctx, mid, True, obs_dev.data_ptr(), pts_dev.data_ptr(), P.width, P.height, pts_period=N_PTS,
                             device_ptrs=True, n_blocks=n_views, obs_per_block=N_PTS, n_blocks_global=n_views * world)
        solver.set_params(P.focal_init, P.pp_init, P.dist_init, P.poses_init)
        solve_exactly(solver, P, max(W, 3), fixed)                   # warm-up iterations (untimed)
        solver.set_params(P.focal_init, P.pp_init, P.dist_init, P.poses_init)
        if launches0 is None:
            launches0 = ctx.kernel_launches()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        summ = solve_exactly(solver, P, K, fixed)
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        iters = summ["iterations"] - 1
        assert iters == K, f"{model}: ran {iters} iterations instead of {K} ({summ})"
        dev_ms += ms
        jac_evals += summ["jacobian_evaluations"]
        detail[model] = {"ms_per_iter": round(ms / K, 4), "jacobian_evals": summ["jacobian_evaluations"],
                         "accepted": summ["num_successful"] - 1}
        # -------- dominant kernel alone (roofline): the Jacobian / normal-equation build
        acc_ms = time_accumulate(ctx, solver, stream, reps=5)
        acc_ms_total += acc_ms
        acc_flops_total += flops_per_obs(model) * float(summ["n_residual_blocks"]) / max(world, 1)
        detail[model]["accumulate_ms"] = round(acc_ms, 4)
        solver.close()
        del obs_dev
        # -------- e2e: host buffers through the public C ABI, copies inside the timed region.  One solve takes ~12 ms of
        # wall time, so a single host hiccup (another tenant, the clock sampler) can be 10x the signal: the median of
        # three identical solves is reported.
        e2e_runs = []
        for _rep in range(3):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0.record(stream)
            s2 = lm.LmSolver(ctx, mid, True, obs_pin.numpy(), P.target, P.width, P.height, n_blocks_global=n_views * world)
            s2.set_params(P.focal_init, P.pp_init, P.dist_init, poses_pin.numpy())
            summ2 = solve_exactly(s2, P, K, fixed, poses_pin.numpy())
            out = s2.get_params()
            e1.record(stream)
            torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms2], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms2 = float(t.item())
            e2e_runs.append(ms2)
            s2.close()
        ms2 = float(np.median(e2e_runs))
        e2e_ms += ms2
        detail[model]["e2e_ms_per_iter"] = round(ms2 / K, 4)
        h2d_bytes += P.obs.nbytes + P.poses_init.nbytes + P.target.nbytes + 8 * 12
        d2h_bytes += out[3].nbytes + 8 * 12 + K * 8 * 261
    clocks = sampler.stop()
    launches = ctx.kernel_launches() - (launches0 or 0)

    total_steps = K * len(models)
    obs_scale = (n_views * N_PTS * world) / OBS_10M
    value = total_steps / (dev_ms * 1e-3) * obs_scale
    e2e_value = total_steps / (e2e_ms * 1e-3) * obs_scale
    achieved = acc_flops_total / (acc_ms_total * 1e-3) / 1e12
    line = {
        "metric": "LM iters/s (10M obs)", "value": round(value, 3), "unit": "LM iters/s (10M obs)",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": round(dev_ms / total_steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg5 LM-only sweep: %d views x %d pts = %d observations per GPU, 7 dist_models (mono focal), "
                               "stage-3 pattern (focal+pp+dist+poses), Cauchy loss; value aggregates the 7 models" % (n_views, N_PTS, n_views * N_PTS),
                   "models": list(models), "per_model": detail, "l2": "inputs (80 MB obs + 89 MB strips per model) exceed L2 only partly; "
                   "each iteration rewrites 89 MB of per-view strips + 55 MB Schur records between reads",
                   "jacobian_evaluations": jac_evals},
        "e2e": {"value": round(e2e_value, 3), "unit": "LM iters/s (10M obs)",
                "h2d_bytes_per_step": int(h2d_bytes / total_steps), "d2h_bytes_per_step": int(d2h_bytes / total_steps)},
        "gpu_launches": int(launches),
        "exchange": exchange,
        "parity_note": "B8 (ceres::Solve rules) is checked against a RESTATED Ceres (oracle/lm_oracle.cpp); the 14 residual functors and the "
                       "SE3 manifold are pinned to the reference's own headers compiled unmodified (oracle/lm_ref)",
        "clocks": clocks,
        "roofline": {"bound": "fp64", "achieved": round(achieved, 3), "peak": round(peak_fp64, 3), "unit": "TFLOP/s",
                     "frac": round(achieved / peak_fp64, 4) if peak_fp64 else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of ONE radtan5 launch (10,000,080 obs), ncu --set full:
                     # profiles/ncu_lm_r01_final.txt; algorithmic: 80.0 MB obs + 3.9 MB poses read, 35.6 MB strips written
                     "traffic": measured_traffic().get("lm_accumulate_radtan5_dram_bytes_per_launch"),
                     "traffic_launch": "lm_accumulate_kernel<radtan5,mono> on 10,000,080 observations; " + str(measured_traffic().get("source", "no ncu capture in profiles/")),
                     "kernel": "lm_accumulate_kernel", "dmma_peak_tflops": round(peak_dmma, 3), "peak_source": "DFMA microbenchmark measured in this run "
                     "(MEASURED_PEAKS.json has no FP64 figure; hbm_gbs there = %s)" % measured_peaks().get("hbm_gbs")},
    }
    if not args.no_detect:
        line["detect"] = run_detect(ctx, args, rank, world, stream)
        line["gpu_launches"] += line["detect"]["gpu_launches"]
        line["detect_chessboard"] = run_detect_chessboard(ctx, args, rank, world, stream, "cfg1")
        line["detect_chessboard_stereo"] = run_detect_chessboard(ctx, args, rank, world, stream, "cfg2")
        line["gpu_launches"] += line["detect_chessboard"]["gpu_launches"] + line["detect_chessboard_stereo"]["gpu_launches"]
    if not args.no_rig:
        rig = run_cfg4(ctx, args, rank, world, stream)
        if rank == 0:
            line["cfg4"] = rig
            line["gpu_launches"] += rig["gpu_launches"]
    if rank == 0:
        if not args.no_cpu and world >= 1:
            v, ms_cpu, cb = run_cpu(argparse.Namespace(steps=min(K, 10), warmup=1), ("radtan5",), args.cpu_views, False)
            line["cpu_baseline"] = cb
        # compact recap LAST: whoever keeps only the tail of a long line still sees every per-N number
        head = {"n_gpus": world, "lm_iters_s": line["value"], "lm_e2e_iters_s": line["e2e"]["value"], "lm_roofline_frac": line["roofline"]["frac"]}
        if "detect" in line:
            head.update(aprilgrid_img_s=line["detect"]["value"], aprilgrid_e2e_img_s=line["detect"]["e2e"]["value"],
                        chessboard_cfg1_img_s=line["detect_chessboard"]["value"], chessboard_cfg2_img_s=line["detect_chessboard_stereo"]["value"])
        if line.get("cfg4"):
            head.update(cfg4_wall_s=line["cfg4"]["value"], cfg4_frames=line["cfg4"]["frames_total"], cfg4_speedup_vs_cpu=line["cfg4"].get("speedup_vs_cpu"))
        line["headline"] = head
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


# ------------------------------------------------------------------------------------------- detection
DET_W, DET_H, DET_ROWS, DET_COLS = 2448, 2048, 6, 6


def run_detect(ctx, args, rank, world, stream):
    """cfg3-style workload: 6x6 AprilGrid (36h11) views 2448x2048 rendered into HBM, radtan8 camera.
    Returns the `detect` object of the JSON line (images/s detect+subpix, HBM roofline, reference CPU)."""
    import torch
    import torch.distributed as dist

    from ezcalib_b200 import detect, synth
    n = args.det_images
    model = "radtan8"
    mid = synth.MODEL_ID[model]
    d = synth.GT_DIST[model]
    f, cx, cy = 2000.0, DET_W / 2 + 3.0, DET_H / 2 - 4.0
    tgt = synth.aprilgrid_target(DET_ROWS, DET_COLS, 0.06, 0.3)
    rng = np.random.Generator(np.random.Philox(key=synth.MASTER_SEED + 1000 + rank))
    poses = synth.sample_poses(rng, n, tgt, model, f, cx, cy, d, DET_W, DET_H, margin=30, tilt_sigma=0.3)
    cams = np.tile(np.array([f, f, cx, cy] + list(d)), (n, 1))
    imgs = detect.render_targets(ctx, "aprilgrid", DET_ROWS, DET_COLS, 0.06, 0.3, mid, DET_W, DET_H, cams, poses,
                                 seed=synth.MASTER_SEED + rank)
    for _ in range(max(args.warmup, 3) if n <= 64 else 3):
        succ, corners, ndet = detect.detect_aprilgrid(ctx, imgs, DET_ROWS, DET_COLS)
    reps = max(1, args.det_reps)
    l0 = ctx.kernel_launches()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        succ, corners, ndet = detect.detect_aprilgrid(ctx, imgs, DET_ROWS, DET_COLS)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = ctx.kernel_launches() - l0
    # e2e: host (pinned) images -> upload -> detect -> results on the host
    host = imgs.download(0, n)
    pin = torch.from_numpy(host).pin_memory()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    runs = []
    for _rep in range(3):   # median of three identical host-buffer runs (one run is ~0.1 s of wall time: host hiccups show)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        im2 = detect.Images(ctx, pin.numpy(), async_upload=True)   # copies on a second stream overlap detection of earlier sub-batches
        s2, c2, n2 = detect.detect_aprilgrid(ctx, im2, DET_ROWS, DET_COLS)
        e1.record(stream)
        torch.cuda.synchronize()
        runs.append(e0.elapsed_time(e1))
        im2.close()
        assert np.array_equal(c2, corners) and np.array_equal(n2, ndet)
    ms2 = float(np.median(runs))
    # what the PCIe link of this box delivers for the same pinned buffer (context for e2e: it is transfer bound)
    dst = torch.empty(pin.numel(), dtype=torch.uint8, device="cuda")
    dst.copy_(pin.view(-1), non_blocking=True)
    torch.cuda.synchronize()
    e0.record(stream)
    dst.copy_(pin.view(-1), non_blocking=True)
    e1.record(stream)
    torch.cuda.synchronize()
    h2d_gbs = pin.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9
    del dst
    if world > 1:
        t = torch.tensor([ms, ms2], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms2 = float(t[0].item()), float(t[1].item())
    value = n * reps * world / (ms * 1e-3)
    e2e_value = n * world / (ms2 * 1e-3)
    alg_bytes = DET_W * DET_H + 8 * 4 * DET_ROWS * DET_COLS
    peak = measured_peaks().get("hbm_gbs", 6650.0)
    achieved = value / world * alg_bytes / 1e9
    out = {"metric": "images/s detect+subpix", "value": round(value, 2), "unit": "images/s", "ms_per_image": round(ms / (n * reps), 4),
           "config": {"workload": "cfg3-style: %d views %dx%d per GPU, 6x6 AprilGrid 36h11, radtan8, rendered into HBM (inputs %.1f GB > L2)" % (n, DET_W, DET_H, n * DET_W * DET_H / 1e9),
                      "found": int(succ.sum()), "mean_detections": float(ndet.mean())},
           "e2e": {"value": round(e2e_value, 2), "unit": "images/s", "h2d_bytes_per_step": int(DET_W * DET_H), "d2h_bytes_per_step": int(8 * 144 + 5),
                   "h2d_link_gbs": round(h2d_gbs, 2), "h2d_bound_images_per_s": round(h2d_gbs * 1e9 / (DET_W * DET_H) * world, 1)},
           "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 5),
                        "traffic": measured_traffic().get("aprilgrid_pipeline_dram_bytes_per_image"),
                        "kernel": "whole detection pipeline (per-kernel shares: profiles/launch_shares_r02.txt)",
                        "top_kernel_traffic": {"kernel": "at_merge_kernel", "dram_bytes_per_image": measured_traffic().get("at_merge_dram_bytes_per_image"),
                                               "algorithmic_bytes_per_image": alg_bytes, "source": measured_traffic().get("source")},
                        "peak_source": "MEASURED_PEAKS.json hbm_gbs" if measured_peaks().get("hbm_gbs") else "fallback 6650"}}
    if rank == 0 and not args.no_cpu:
        from oracle import apriltag_oracle as A
        k = min(n, args.det_cpu_images)
        A.detect_target(host[0], DET_ROWS, DET_COLS)
        t0 = time.perf_counter()
        same = True
        for i in range(k):
            ok, c, nd, _ = A.detect_target(host[i], DET_ROWS, DET_COLS)
            same = same and bool(ok) == bool(succ[i]) and int(nd) == int(ndet[i]) and np.array_equal(c[:, 0] < 0, corners[i][:, 0] < 0) \
                and float(np.abs(c - corners[i]).max()) <= 0.01
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": round(k / dt, 3), "unit": "images/s", "cores": host_cores(), "kind": "reference",
                               "sample": "%d of the same images: the reference's own extractTags (oracle/_ref, OpenMP) + cv2.cornerSubPix, serial image loop like ezcalibrator.cpp:40-59" % k,
                               "parity_on_sample": bool(same)}
    imgs.close()
    return out


# ------------------------------------------------------------------------------------------- cfg4: the 8-camera rig
RIG_W, RIG_H, RIG_CAMS = 3840, 2160, 8


def rig_ground_truth():
    """8 cameras around cam0 (all overlap cam0): per-camera intrinsics, radtan8 distortion, T_cj_c0."""
    from ezcalib_b200 import synth
    rng = np.random.Generator(np.random.Philox(key=synth.MASTER_SEED + 4000))
    d0 = np.array(synth.GT_DIST["radtan8"])
    cams, ext = [], []
    for c in range(RIG_CAMS):
        f = 2750.0 * (1 + 0.01 * rng.standard_normal())
        cams.append(dict(f=f, cx=RIG_W / 2 + 6 * rng.standard_normal(), cy=RIG_H / 2 + 6 * rng.standard_normal(),
                         d=d0 * (1 + 0.05 * rng.standard_normal(len(d0)))))
        if c == 0:
            ext.append(np.array([0, 0, 0, 1.0, 0, 0, 0]))
        else:
            ext.append(np.concatenate([synth.quat_from_rotvec(np.deg2rad(1.5) * rng.standard_normal(3)),
                                       np.array([0.06, 0.04, 0.01]) * rng.standard_normal(3)]))
    return cams, ext


def run_cfg4(ctx, args, rank, world, stream):
    """BASELINE.json configs[3]: 8 cameras x n synchronised 3840x2160 frames (6x6 AprilGrid, radtan8), camera c on GPU c % N:
    detect -> P3P-RANSAC -> 3-stage solve per camera, cam0's poses broadcast, joint extrinsic solve with scalar all-reduces
    (apps/ezcalib.cpp:205-218, src/ezcalibrator.cpp:20-67, :631-844).  Frames are rendered into HBM in chunks outside the
    timed regions (they stand for cv::imread, which SURVEY 8d excludes on both sides)."""
    import torch
    import torch.distributed as dist

    from ezcalib_b200 import detect, ezcalibrator as E, synth
    from ezcalib_b200 import dist as ezdist
    n = args.rig_frames
    model, mono, fov = "radtan8", True, 70.0
    mid = synth.MODEL_ID[model]
    tgt = synth.aprilgrid_target(DET_ROWS, DET_COLS, 0.06, 0.3)
    gt_cams, gt_ext = rig_ground_truth()
    rng = np.random.Generator(np.random.Philox(key=synth.MASTER_SEED + 4001))
    c0 = gt_cams[0]
    poses0 = synth.sample_poses(rng, n, tgt, model, c0["f"], c0["cx"], c0["cy"], c0["d"], RIG_W, RIG_H, margin=260, tilt_sigma=0.3)
    names = ["%06d.png" % i for i in range(n)]
    mine = [c for c in range(RIG_CAMS) if c % world == rank]
    comm = ezdist.TorchComm() if world > 1 else E.SingleProcessComm()
    chunk = max(1, min(n, args.rig_chunk))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    det_ms, n_imgs, launches0 = 0.0, 0, ctx.kernel_launches()
    timings, local, sample = {}, {}, None
    for c in mine:
        g = gt_cams[c]
        poses_c = np.stack([synth.se3_mul(gt_ext[c], p) for p in poses0])
        cam12 = np.array([g["f"], g["f"], g["cx"], g["cy"]] + list(g["d"]))
        found = np.zeros(n, bool)
        corners = np.zeros((n, len(tgt), 2), np.float32)
        for b in range(0, n, chunk):
            m = min(chunk, n - b)
            imgs = detect.render_targets(ctx, "aprilgrid", DET_ROWS, DET_COLS, 0.06, 0.3, mid, RIG_W, RIG_H, np.tile(cam12, (m, 1)),
                                         poses_c[b:b + m], seed=synth.MASTER_SEED + 100 * c + b)
            if c == mine[0] and b == 0:
                detect.detect_aprilgrid(ctx, imgs, DET_ROWS, DET_COLS)   # warm-up (pools, first-launch costs), untimed
                if rank == 0 and not args.no_cpu:
                    k = min(m, args.rig_cpu_frames)
                    sample = dict(images=imgs.download(0, k), cam=c)
            torch.cuda.synchronize()
            e0.record(stream)
            f_, c_, _ = detect.detect_aprilgrid(ctx, imgs, DET_ROWS, DET_COLS)
            e1.record(stream)
            torch.cuda.synchronize()
            det_ms += e0.elapsed_time(e1)
            n_imgs += m
            found[b:b + m] = f_
            corners[b:b + m] = c_
            imgs.close()
        cam, processed, _ = E.calibrate_camera_from_corners(ctx, found, corners, names, tgt, model, mono, fov, RIG_W, RIG_H, timings)
        local[c] = cam
        if sample is not None and sample["cam"] == c:
            sample.update(found=found[:len(sample["images"])].copy(), corners=corners[:len(sample["images"])].copy())
    if world > 1:
        dist.barrier()
    ext, msum = E.run_rig_extrinsics(ctx, local, RIG_CAMS, tgt, comm, timings)
    torch.cuda.synchronize()
    launches = ctx.kernel_launches() - launches0
    stage = np.array([det_ms * 1e-3, timings.get("p3p_s", 0.0), timings.get("solve_s", 0.0),
                      timings.get("multicam_setup_s", 0.0) + timings.get("multicam_solve_s", 0.0)])
    wall = float(stage.sum())
    if world > 1:
        t = torch.tensor(list(stage) + [wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        stage, wall = t[:4].cpu().numpy(), float(t[4].item())
    # sanity against the ground truth of the renders (calibration quality, not parity)
    stats = []
    for c in mine:
        cam, g = local[c], gt_cams[c]
        rm = np.array([f.rmse_err for f in cam.frames])
        e = dict(cam=c, frames=len(cam.frames), focal_rel_err=abs(cam.focal[0] - g["f"]) / g["f"], median_rmse_px=float(np.median(rm)),
                 frames_rmse_gt_1px=int((rm > 1.0).sum()))
        if c >= 1:
            e["baseline_err_mm"] = float(1e3 * np.linalg.norm(ext[c][4:] - gt_ext[c][4:]))
        stats.append(e)
    allstats = comm.allgather(stats)
    out = None
    if rank == 0:
        flat = sorted((e for part in allstats for e in part), key=lambda e: e["cam"])
        total_frames = RIG_CAMS * n
        out = {"metric": "calibration wall-time, 8-camera rig", "value": round(wall, 4), "unit": "s", "higher_is_better": False,
               "config": {"workload": "cfg4: %d cameras x %d synchronised %dx%d frames, 6x6 AprilGrid 36h11, radtan8 mono focal; camera c on GPU c %% %d; "
                                      "frames rendered into HBM in chunks of %d outside the timed regions" % (RIG_CAMS, n, RIG_W, RIG_H, world, chunk)},
               "n_gpus": world, "frames_total": total_frames,
               "stage_s": {"detect": round(float(stage[0]), 4), "p3p": round(float(stage[1]), 4), "mono_solve": round(float(stage[2]), 4),
                           "multicam": round(float(stage[3]), 4)},
               "detect_images_per_s": round(n_imgs * world / max(float(stage[0]), 1e-9), 1),
               "multicam": {"iterations": msum["iterations"], "termination": msum["termination_name"], "n_good_pairs": msum["n_good_pairs"],
                            "setup_s_rank0": round(timings.get("multicam_setup_s", 0.0), 4), "solve_s_rank0": round(timings.get("multicam_solve_s", 0.0), 4)},
               "per_camera": flat, "gpu_launches": int(launches),
               "sane": bool(all(e["focal_rel_err"] < 5e-3 and e["median_rmse_px"] < 0.5 for e in flat) and
                            all(e.get("baseline_err_mm", 0.0) < 5.0 for e in flat))}
        if sample is not None:
            out["cpu_baseline"] = cfg4_cpu_baseline(sample, tgt, model, mono, fov, n, local[sample["cam"]])
            if out["cpu_baseline"].get("wall_s_extrapolated"):
                out["speedup_vs_cpu"] = round(out["cpu_baseline"]["wall_s_extrapolated"] / wall, 1)
    return out


def cfg4_cpu_baseline(sample, tgt, model, mono, fov, n_frames, cam_gpu):
    """The reference's serial flow (ezcalibrator.cpp:40-59, apps/ezcalib.cpp:205-218) timed on a stated sample of the same
    frames and extrapolated linearly to the whole rig: extractTags (the reference's own sources, oracle/_ref, OpenMP) +
    cv2.cornerSubPix per frame, cv2.solvePnPRansac per frame, the restated-Ceres 3-stage solve on the sampled views."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "p3p_harness"))
    import p3p_cv2_ref as PR
    from ezcalib_b200 import ezcalibrator as E, synth
    from oracle import apriltag_oracle as A
    from oracle import lm_oracle as O
    imgs = sample["images"]
    k = len(imgs)
    focal0, pp0 = E.setup_initial_calibration(RIG_W, RIG_H, fov, mono)
    A.detect_target(imgs[0], DET_ROWS, DET_COLS)
    t0 = time.perf_counter()
    det_same, obs, poses = True, [], []
    for i in range(k):
        ok, c, nd, _ = A.detect_target(imgs[i], DET_ROWS, DET_COLS)
        det_same = det_same and bool(ok) == bool(sample["found"][i]) and np.array_equal(c[:, 0] < 0, sample["corners"][i][:, 0] < 0) \
            and float(np.abs(c - sample["corners"][i]).max()) <= 0.01
        if ok:
            obs.append(c.astype(np.float32))
    t1 = time.perf_counter()
    p3p_same = True
    by_name = {f.img_name: f for f in cam_gpu.frames}
    for c in obs:
        okp, pose, _, _ = PR.compute_p3p_ransac_cv2(c, tgt, focal0, pp0, RIG_W, RIG_H)
        if okp:
            poses.append(pose)
    t2 = time.perf_counter()
    n_lm = len(poses)
    lm_s = 0.0
    if n_lm >= 4:
        O.solve_mono(synth.MODEL_ID[model], mono, np.stack(obs[:n_lm]), tgt, RIG_W, RIG_H, focal0, pp0, np.zeros(8), np.stack(poses), threads=host_cores())
        lm_s = time.perf_counter() - t2
    per_frame = (t2 - t0) / k + (lm_s / n_lm if n_lm else 0.0)
    del by_name, p3p_same
    return {"value": round(k / (t2 - t0 + lm_s), 4), "unit": "frames/s (detect + P3P + solve)", "cores": host_cores(), "kind": "reference",
            "sample": "%d frames of camera %d: the reference's own extractTags (oracle/_ref, OpenMP) + cv2.cornerSubPix, cv2.solvePnPRansac, "
                      "restated-Ceres 3-stage solve on those views; serial frame loop like ezcalibrator.cpp:40-59" % (k, sample["cam"]),
            "detect_s_per_frame": round((t1 - t0) / k, 4), "p3p_s_per_frame": round((t2 - t1) / max(len(obs), 1), 6),
            "solve_s_per_view": round(lm_s / n_lm, 6) if n_lm else None,
            "wall_s_extrapolated": round(per_frame * RIG_CAMS * n_frames, 1),
            "extrapolation": "per-frame cost x %d frames (the multi-camera extrinsic solve is not included on the CPU side)" % (RIG_CAMS * n_frames),
            "detect_parity_on_sample": bool(det_same)}


CB_CONFIGS = {
    # BASELINE.json configs[0]: calib_chessboard_mono_config.yaml geometry (more views than the reference's 50 so that
    # the batch exceeds L2 and the timing is not launch noise)
    "cfg1": dict(W=1280, H=1024, model="radtan5", f=900.0, cx=645.0, cy=505.0, margin=90, label="cfg1-style (mono, rad-tan 5)"),
    # BASELINE.json configs[1]: calib_chessboard_stereo_config.yaml geometry: 2 cameras x 200 views 1920x1080, Kannala-Brandt
    "cfg2": dict(W=1920, H=1080, model="kannalabrandt", f=900.0, cx=965.0, cy=533.0, margin=90, label="cfg2-style (stereo pair = 2 x 200 views, Kannala-Brandt 4)"),
}


def run_detect_chessboard(ctx, args, rank, world, stream, cfg="cfg1"):
    """9x6 chessboard (10x7 squares) views rendered into HBM; cfg1 = mono 1280x1024 rad-tan 5, cfg2 = stereo 1920x1080 KB4."""
    import torch

    from ezcalib_b200 import detect, synth
    C = CB_CONFIGS[cfg]
    n, W, H, model = args.cb_images, C["W"], C["H"], C["model"]
    mid = synth.MODEL_ID[model]
    d = synth.GT_DIST[model]
    f, cx, cy = C["f"], C["cx"], C["cy"]
    tgt = synth.chessboard_target(7, 10, 0.05)
    rng = np.random.Generator(np.random.Philox(key=synth.MASTER_SEED + 2000 + rank))
    poses = synth.sample_poses(rng, n, tgt, model, f, cx, cy, d, W, H, margin=C["margin"], tilt_sigma=0.3)
    cams = np.tile(np.array([f, f, cx, cy] + list(d) + [0.0] * (8 - len(d))), (n, 1))
    imgs = detect.render_targets(ctx, "chessboard", 7, 10, 0.05, 0.0, mid, W, H, cams, poses, seed=synth.MASTER_SEED + 7 + rank)
    # Warm-up until the call time is steady (at least 3 calls, at most 4 s): this detector is a long chain of short launches
    # and host round trips, and for about a second after the 16-thread cv2 baseline of the previous block its calls take
    # up to twice as long (the pool's worker threads are still spinning on the host cores).
    t_warm = time.perf_counter()
    warm = []
    while True:
        t0 = time.perf_counter()
        found, corners = detect.detect_chessboard(ctx, imgs, 7, 10)
        warm.append(time.perf_counter() - t0)
        if len(warm) >= 3 and (max(warm[-3:]) <= 1.05 * min(warm[-3:]) or time.perf_counter() - t_warm > 4.0):
            break
    l0 = ctx.kernel_launches()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # seven timed calls, each bracketed by events; the MEDIAN call is reported (a call is 30-60 ms of many short launches
    # and host round trips, and in a full bench run single calls take up to twice that -- the mean of three consecutive
    # calls moved by 2x from run to run)
    reps = 1
    calls_ms = []
    for _ in range(7):
        torch.cuda.synchronize()
        e0.record(stream)
        found, corners = detect.detect_chessboard(ctx, imgs, 7, 10)
        e1.record(stream)
        torch.cuda.synchronize()
        calls_ms.append(e0.elapsed_time(e1))
    ms = float(np.median(calls_ms))
    # e2e: pinned host images -> (async) upload -> detect -> corners on the host
    host_all = imgs.download(0, n)
    pin = torch.from_numpy(host_all).pin_memory()
    torch.cuda.synchronize()
    runs = []
    for _rep in range(3):   # median of three identical host-buffer runs
        torch.cuda.synchronize()
        e0.record(stream)
        im2 = detect.Images(ctx, pin.numpy(), async_upload=True)
        f2, c2 = detect.detect_chessboard(ctx, im2, 7, 10)
        e1.record(stream)
        torch.cuda.synchronize()
        runs.append(e0.elapsed_time(e1))
        im2.close()
        assert np.array_equal(f2, found) and np.array_equal(c2, corners)
    ms2 = float(np.median(runs))
    if world > 1:   # weak scaling: n views per rank, slowest rank decides
        import torch.distributed as dist
        t = torch.tensor([ms, ms2], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms2 = float(t[0].item()), float(t[1].item())
    value = n * reps * world / (ms * 1e-3)
    alg_bytes = W * H + 8 * 54
    peak = measured_peaks().get("hbm_gbs", 6650.0)
    out = {"metric": "images/s detect+subpix (chessboard)", "value": round(value, 2), "unit": "images/s",
           "config": {"workload": "%s: %d views %dx%d, 9x6 inner corners, rendered into HBM (inputs %.2f GB > L2)" % (C["label"], n, W, H, n * W * H / 1e9),
                      "found": int(found.sum()), "timed_calls_ms": [round(r, 2) for r in calls_ms], "reported": "median call", "warmup_calls": len(warm)},
           "e2e": {"value": round(n * world / (ms2 * 1e-3), 2), "unit": "images/s", "h2d_bytes_per_step": int(W * H), "d2h_bytes_per_step": 8 * 54 + 1},
           "gpu_launches": int(ctx.kernel_launches() - l0),
           "roofline": {"bound": "hbm", "achieved": round(value / world * alg_bytes / 1e9, 3), "peak": peak, "unit": "GB/s",
                        "frac": round(value / world * alg_bytes / 1e9 / peak, 6), "traffic": None}}
    if rank == 0 and not args.no_cpu:
        import cv2
        flags = cv2.CALIB_CB_ADAPTIVE_THRESH + cv2.CALIB_CB_NORMALIZE_IMAGE + cv2.CALIB_CB_FILTER_QUADS + cv2.CALIB_CB_FAST_CHECK
        crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 100, 0.05)
        # sample with the SAME mix as the batch: up to 46 views the GPU found + every view it did not find (cv2 pays for
        # those too: all 8 + 48 attempts run), weighted back to the batch composition
        idx_found = [i for i in range(n) if found[i]][:46]
        idx_fail = [i for i in range(n) if not found[i]][:6]
        worst, flags_equal = 0.0, True
        t_found = t_fail = 0.0
        for i in idx_found + idx_fail:
            t0 = time.perf_counter()
            ok, c = cv2.findChessboardCorners(host_all[i], (9, 6), flags=flags)
            if ok:
                c = cv2.cornerSubPix(host_all[i], c, (5, 5), (-1, -1), crit).reshape(-1, 2)
            dt = time.perf_counter() - t0
            if found[i]:
                t_found += dt
            else:
                t_fail += dt
            flags_equal = flags_equal and bool(ok) == bool(found[i])
            if ok and found[i]:
                worst = max(worst, float(np.abs(c - corners[i]).max()))
        # parity against OpenCV's generic code path (what the kernels restate; IPP's sampler moves weak corners): untimed
        ipp_was = cv2.ipp.useIPP()
        cv2.ipp.setUseIPP(False)
        worst_generic = 0.0
        for i in idx_found[:12]:
            ok, c = cv2.findChessboardCorners(host_all[i], (9, 6), flags=flags)
            if ok:
                c = cv2.cornerSubPix(host_all[i], c, (5, 5), (-1, -1), crit).reshape(-1, 2)
                worst_generic = max(worst_generic, float(np.abs(c - corners[i]).max()))
        cv2.ipp.setUseIPP(ipp_was)
        n_found, n_fail = int(found.sum()), int(n - found.sum())
        est = (t_found / max(len(idx_found), 1)) * n_found + (t_fail / max(len(idx_fail), 1)) * n_fail
        out["cpu_baseline"] = {"value": round(n / est, 3), "unit": "images/s", "cores": host_cores(), "kind": "reference",
                               "sample": "%d found + %d not-found views of the batch, cv2.findChessboardCorners + cv2.cornerSubPix with the reference's flags "
                                         "(cv2 %s, %d threads), weighted to the batch mix (%d found / %d not found); mean %.1f ms per found view, %.1f ms per "
                                         "not-found view" % (len(idx_found), len(idx_fail), cv2.__version__, cv2.getNumThreads(), n_found, n_fail,
                                                             1e3 * t_found / max(len(idx_found), 1), 1e3 * t_fail / max(len(idx_fail), 1)),
                               "found_flags_equal": bool(flags_equal), "max_corner_diff_px": round(worst, 6),
                               "max_corner_diff_px_without_ipp": round(worst_generic, 6)}
    imgs.close()
    return out


def time_accumulate(ctx, solver, stream, reps=5) -> float:
    """Average duration (ms) of the Jacobian/normal-equation build alone, CUDA events on the launch stream."""
    import torch
    solver.normal_equations(1, 1, 1)  # warm
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    from ezcalib_b200 import _lib
    import ctypes as C
    best = []
    for _ in range(reps):
        e0.record(stream)
        _lib.check(_lib.lib().ezc_lm_accumulate_only(solver._h, 1, 1, 1), "ezc_lm_accumulate_only")
        e1.record(stream)
        torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1))
    del C
    return float(np.mean(best))


_REAL_STDOUT = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to fd 1 when
    NCCL_DEBUG is set), so fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--models", default="")
    ap.add_argument("--views", type=int, default=N_VIEWS_10M)
    ap.add_argument("--cpu-views", type=int, default=20000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-detect", action="store_true")
    ap.add_argument("--det-images", type=int, default=500)
    ap.add_argument("--det-reps", type=int, default=2)
    ap.add_argument("--cb-images", type=int, default=400)
    ap.add_argument("--det-cpu-images", type=int, default=24)
    ap.add_argument("--no-rig", action="store_true")
    ap.add_argument("--rig-frames", type=int, default=2000)
    ap.add_argument("--rig-chunk", type=int, default=500)
    ap.add_argument("--rig-cpu-frames", type=int, default=6)
    ap.add_argument("--only-rig", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        if rank != 0:
            return
        models = BENCH_MODELS if not args.models else tuple(args.models.split(","))
        # SAME configuration as our arm: every model on the full 69,445-view / 10,000,080-observation problem.  A CPU iteration
        # of that size takes ~1.3 s per model on 16 cores, so a step budget bounds the run: at most 2 timed iterations (and 1
        # warm-up iteration) per model however large --steps is -- ~1 minute in total.
        a2 = argparse.Namespace(steps=max(1, min(args.steps, 2)), warmup=1)
        views = args.views if args.cpu_views == 20000 else args.cpu_views
        value, ms, cb = run_cpu(a2, models, views, True)
        line = {"impl": "reference", "metric": "LM iters/s (10M obs)", "value": round(value, 6), "unit": "LM iters/s (10M obs)",
                "n_gpus": args.gpus, "steps": a2.steps, "warmup": a2.warmup, "steps_requested": args.steps, "warmup_requested": args.warmup,
                "ms_per_step": round(ms, 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "cfg5 LM-only sweep, same problem as the GPU arm (bounded in iterations, not in size): " + cb["sample"],
                           "note": "Ceres 2.2.0 is not installable here (no source, no network): this is the restated-Ceres oracle port with all host threads"},
                "cpu_baseline": cb,
                "e2e": {"value": round(value, 6), "unit": "LM iters/s (10M obs)", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return
    run_gpu(args)


if __name__ == "__main__":
    main()
