#!/usr/bin/env python
"""bench.py -- headline benchmark of the transient path-tracing hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): M-paths/sec on the Cornell-box transient scene, 512x512 px x 1024 time bins,
256 spp, max_depth 8 (config C2).  One "step" = one full render of that configuration:
  value  -- device-resident: tr_render into the raw (H,W,T,RGBW) film in HBM, CUDA-event timed.
  e2e    -- through the host API with HOST buffers: scene upload (H2D) + render + film copied to pinned
            host memory (D2H) inside the timed region (tr_scene_create + tr_render_to_host).
N > 1 (torchrun): pixel tiles are dealt round-robin to the ranks, per-GPU work is fixed (spp = 256*N,
weak scaling), and one NCCL reduce of the film to rank 0 closes every step (inside the timed region).

--impl reference: the reference's own implementation (mitsuba + drjit) cannot be installed in this image
(no wheel, no network), so the arm times the CPU oracle port of the same path (oracle/, scalar C++,
all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C2 = dict(width=512, height=512, temporal_bins=1024, spp=256, max_depth=8)
METRIC = "M-paths/sec (Cornell box transient_path, 512x512 px x 1024 t-bins, 256 spp, max_depth 8)"
UNIT = "M-paths/s"


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch():
    """dram bytes per launch of the render kernel at the bench workload, from the committed ncu summary."""
    p = os.path.join(ROOT, "profiles", "ncu_render_kernel_summary.json")
    try:
        with open(p) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_rate(mitr, oracle, width, height, bins, spp, max_depth, threads=0):
    """Time one oracle render; returns (M-paths/s, seconds, counters)."""
    scene = mitr.load_dict(mitr.cornell_box_config(width, height, bins, spp, max_depth))
    t0 = time.perf_counter()
    r = oracle.render(scene, scene.integrator(), seed=0, spp=spp, n_threads=threads)
    dt = time.perf_counter() - t0
    return width * height * spp / dt / 1e6, dt, r["counters"]


def pick_cpu_sample(mitr, oracle, target_s):
    """Largest square sub-resolution of C2 (same spp, bins, depth) the oracle renders in ~target_s."""
    oracle_rate(mitr, oracle, 32, 32, C2["temporal_bins"], 16, C2["max_depth"])            # warm-up (C1-sized)
    rate, _, _ = oracle_rate(mitr, oracle, 32, 32, C2["temporal_bins"], C2["spp"], C2["max_depth"])
    paths = rate * 1e6 * target_s
    side = int((paths / C2["spp"]) ** 0.5) // 16 * 16
    return max(32, min(C2["width"], side))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import mitransient_b200 as mitr
    from oracle import oracle
    oracle.build()
    cores = len(os.sched_getaffinity(0))
    side = pick_cpu_sample(mitr, oracle, target_s=4.0)
    for _ in range(args.warmup):
        oracle_rate(mitr, oracle, side, side, C2["temporal_bins"], C2["spp"], C2["max_depth"])
    t_total, paths = 0.0, 0
    for _ in range(args.steps):
        _, dt, _ = oracle_rate(mitr, oracle, side, side, C2["temporal_bins"], C2["spp"], C2["max_depth"])
        t_total += dt; paths += side * side * C2["spp"]
    value = paths / t_total / 1e6
    sample = (f"Cornell box {side}x{side} px x {C2['temporal_bins']} bins x {C2['spp']} spp, max_depth 8 "
              f"({side * side * C2['spp']} paths per step)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2 cornell_box transient_path 512x512x1024 256spp depth8 (bounded CPU sample)",
                   "note": "mitsuba/drjit not installable here: CPU oracle port of the reference path, scalar C++"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    import mitransient_b200 as mitr
    from mitransient_b200 import _abi, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: mitransient_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = world
    spp = C2["spp"] * n                                        # weak scaling: 2^26 paths per GPU
    W, H, T = C2["width"], C2["height"], C2["temporal_bins"]
    scene = mitr.load_dict(mitr.cornell_box_config(W, H, T, spp, C2["max_depth"]))
    integ = scene.integrator()
    sensor = scene.sensors()[0]
    film_plugin = sensor.film()
    dev = integ.device_scene(scene, sensor)
    lib = _lib.load()
    stream = torch.cuda.current_stream()

    # N > 1: "peer" = every rank flushes its pixel histograms straight into rank 0's film over NVLink peer
    # memory (fused into the kernel, one tiny all-reduce as a fence); "reduce" = private films + one NCCL reduce.
    mode = args.multi_gpu_mode if n > 1 else "single"
    peer = None
    if mode == "peer":
        try:
            from mitransient_b200.distributed import PeerFilm
            peer = PeerFilm((H, W, T, 4), (H, W, 3), dst=0)
            film, steady = peer.film, peer.steady
        except Exception as exc:                               # no P2P / IPC on this box: fall back to the reduce
            if rank == 0:
                print(f"peer film unavailable ({exc}); falling back to NCCL reduce", file=sys.stderr)
            mode, peer = "reduce", None
        flag = torch.tensor([1.0 if mode == "peer" else 0.0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if flag.item() == 0.0:
            mode, peer = "reduce", None
    if peer is None:
        film = torch.zeros((H, W, T, 4), dtype=torch.float32, device="cuda")
        steady = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
    counters = torch.zeros(8, dtype=torch.int64, device="cuda")
    params = integ.make_params(spp, spp, 0, rank=rank, world=n, film_channels=4)
    launches = [0]

    def finish_step():
        if mode == "peer":
            peer.fence()
        elif mode == "reduce":
            dist.reduce(film, dst=0)
            dist.reduce(steady, dst=0)

    def step(with_counters=False):
        if mode == "reduce" and rank == 0:
            film.zero_()                                       # last step's reduce landed in rank 0's film
        _lib.check(lib.tr_render(dev.handle, C.byref(params), film.data_ptr(), steady.data_ptr(),
                                 counters.data_ptr() if with_counters else None, stream.cuda_stream))
        launches[0] += 1
        finish_step()

    # ---- counters (N_v) from one untimed step
    step(with_counters=True)
    torch.cuda.synchronize()
    cnt = counters.cpu().tolist()
    if n > 1:
        ct = counters.clone(); dist.all_reduce(ct); cnt = ct.cpu().tolist()
    paths_total, n_v = cnt[0], cnt[1]
    assert paths_total == W * H * spp, (paths_total, W * H * spp)

    for _ in range(max(0, args.warmup - 1)):
        step()
    clocks = ClockSampler(local_rank)
    if n > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if rank == 0:
        clocks.start()
    launches[0] = 0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    kern = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev[0].record(stream)
    for i in range(args.steps):
        if mode == "reduce" and rank == 0:
            film.zero_()
        kern[i][0].record(stream)
        _lib.check(lib.tr_render(dev.handle, C.byref(params), film.data_ptr(), steady.data_ptr(), None,
                                 stream.cuda_stream))
        kern[i][1].record(stream)
        launches[0] += 1
        finish_step()
        ev[i + 1].record(stream)
    torch.cuda.synchronize()
    if n > 1:
        dist.barrier()
    clk = clocks.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    kern_ms = sum(a.elapsed_time(b) for a, b in kern) / args.steps
    t = torch.tensor([total_ms, kern_ms], dtype=torch.float64, device="cuda")
    kern_ms_per_rank = [kern_ms]
    if n > 1:
        per_rank = [torch.zeros_like(t) for _ in range(n)]
        dist.all_gather(per_rank, t)
        kern_ms_per_rank = [float(x[1]) for x in per_rank]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_ms = t.tolist()
    ms_per_step = total_ms / args.steps
    value = paths_total / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: host buffers, scene upload + render + film to pinned host memory inside the timed region.
    # N > 1: the SAME job as `value` (one film, spp = 256 * N, pixel tiles sharded over the ranks); every rank
    # renders its tiles and moves its share of the developed film to its own pinned host buffer over its own
    # PCIe link (tr_render_to_host with world = N: compact (n_owned, T, 3) output in tr_owned_pixels order).
    e2e = None
    try:
        from mitransient_b200.distributed import owned_pixels
        n_owned = len(owned_pixels(W, H, 16, rank, n)) if n > 1 else W * H
        host_T = torch.empty((n_owned, T, 3), dtype=torch.float32, pin_memory=True)
        host_S = torch.empty((n_owned, 3), dtype=torch.float32, pin_memory=True)
        p_e2e = integ.make_params(spp, spp, 0, film_channels=3, rank=rank, world=n, tile_size=16)
        desc = scene.describe(sensor)
        blob_bytes = dev.info()["blob_bytes"]

        def e2e_step():
            h = C.c_void_p()
            _lib.check(lib.tr_scene_create(C.byref(desc), C.byref(h)))
            try:
                _lib.check(lib.tr_render_to_host(h, C.byref(p_e2e), host_T.data_ptr(), host_S.data_ptr(),
                                                 None, stream.cuda_stream))
            finally:
                lib.tr_scene_destroy(h)
        for _ in range(2):
            e2e_step()
        k_e2e = max(2, min(args.steps, 5))
        if n > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            e2e_step()
            if n > 1:
                dist.barrier()                                  # a step ends when every rank's shard is on the host
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if n > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = te.item()
        nb = torch.tensor([float(host_T.numel() * 4 + host_S.numel() * 4)], dtype=torch.float64, device="cuda")
        if n > 1:
            dist.all_reduce(nb)
        e2e = {"value": W * H * spp * k_e2e / dt / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int((blob_bytes + n_owned * 4) * n),
               "d2h_bytes_per_step": int(nb.item()),
               "steps": k_e2e, "ms_per_step": 1e3 * dt / k_e2e,
               "what": "tr_scene_create + tr_render_to_host: developed (H,W,T,3) film + steady into pinned host "
                       "memory, D2H overlapped with rendering band by band" +
                       (f"; the film is sharded by pixel tiles: each of the {n} ranks renders its tiles and copies its "
                        "share to its own pinned host buffer over its own PCIe link (bytes summed over ranks)" if n > 1 else "")}
        del host_T, host_S
    except Exception as exc:  # keep the headline line even if the host path fails
        e2e = {"value": None, "unit": UNIT, "error": str(exc)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant (only) kernel: render_transient_path_kernel
    peak, peak_src = measured_peak_hbm()
    n_v_rank = n_v / n                                          # per launch (one launch per rank per step)
    b_alg = 128.0 * n_v_rank + 16.0 * (W * H / n) * T + 12.0 * (W * H / n)
    achieved = b_alg / (kern_ms * 1e-3) / 1e9
    traffic = ncu_traffic_per_launch() if n == 1 else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": "render_transient_path_kernel",
                "kernel_ms": kern_ms, "kernel_ms_per_rank": [round(x, 3) for x in kern_ms_per_rank], "algorithmic_bytes": b_alg, "live_path_vertices": n_v_rank,
                "model": "SURVEY 8d: 128 B x N_v + 16 B x H*W*T + 12 B x H*W (wavefront byte model; this kernel "
                         "keeps path state in registers, so its DRAM traffic is ~ the film write only)"}

    # ---- CPU baseline (oracle port) on a bounded sample, rank 0, N = 1 only
    cpu = None
    if n == 1 and not args.no_cpu_baseline:
        try:
            from oracle import oracle
            oracle.build()
            side = pick_cpu_sample(mitr, oracle, target_s=12.0)
            rate, dt, _ = oracle_rate(mitr, oracle, side, side, T, C2["spp"], C2["max_depth"])
            cpu = {"value": rate, "unit": UNIT, "cores": len(os.sched_getaffinity(0)), "kind": "port",
                   "sample": f"Cornell box {side}x{side} px x {T} bins x {C2['spp']} spp, max_depth 8 "
                             f"({side * side * C2['spp']} paths, {dt:.1f} s); CPU oracle restatement (scalar C++), "
                             "NOT Dr.Jit llvm_ad_rgb (mitsuba is not installable here)"}
        except Exception as exc:
            cpu = {"value": None, "unit": UNIT, "error": str(exc)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C2 cornell_box transient_path {W}x{H}x{T} bins, {spp} spp, max_depth 8, seed 0"
                               + (f" (weak scaling: 256 spp x {n} GPUs, pixel tiles 16x16 skewed round-robin, film assembled on "
                                  f"rank 0 inside the step, mode={mode}: " +
                                  ("kernel flushes into rank 0's film over NVLink peer memory + all-reduce fence)"
                                   if mode == "peer" else "NCCL reduce of film+steady to rank 0)") if n > 1 else ""),
                   "paths_per_step": paths_total, "film_bytes": W * H * T * 16,
                   "l2": "no flush needed: each step writes the 4 GiB film (>> 126 MB L2); the scene lives in shared memory",
                   "parallelism": f"pixel-tile x{n}"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches[0], "clocks": clk,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--multi-gpu-mode", default="peer", choices=["peer", "reduce"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
