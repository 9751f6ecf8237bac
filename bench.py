#!/usr/bin/env python
"""bench.py -- headline benchmark of the transient path-tracing hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): M-paths/sec on the Cornell-box transient scene, 512x512 px x 1024 time bins,
256 spp, max_depth 8 (config C2).  One "step" = one full render of that configuration:
  value  -- device-resident: tr_render into the raw (H,W,T,RGBW) film in HBM, CUDA-event timed.
  e2e    -- through the host API with HOST buffers: scene blob upload (H2D) + render + film copied to pinned
            host memory (D2H) inside the timed region (tr_scene_refresh + tr_render_to_host).
  parity -- the SAME C2 render on the CPU oracle (the run that gives `cpu_baseline`): exact work counters and the
            per-time-bin / whole-film relative L1 of the two films (BASELINE.json's second metric).
  extra  -- N = 1 only: the other single-GPU configurations of BASELINE.json (C3 NLOS confocal 256x256x4096,
            C4(i) volumetric Cornell box 256x384x512, the 38-triangle mesh Cornell box), each with kernel time,
            M-paths/s, roofline block, CPU-port rate and a parity block at the full film size.
N > 1 (torchrun): pixel tiles are dealt round-robin to the ranks, per-GPU work is fixed (spp = 256*N,
weak scaling), and the film is assembled on rank 0 inside the timed region (peer stores over NVLink fused into
the kernel, or --multi-gpu-mode reduce: one NCCL reduce).  After the timed loop rank 0 re-renders the same job
alone and compares it with the assembled film (`parity`), and `multi_gpu` records the reduce mode, the strong
scaling case (C2 fixed at 256 spp) and, at N = 8, BASELINE config 5 (4096 spp).

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
N_SM = 148


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_summary(tag):
    """Committed ncu capture of the render kernel at workload `tag` (profiles/ncu_<tag>_summary.json), or None.
    Gives the two numbers a run cannot measure itself: executed warp instructions and DRAM bytes per launch."""
    p = os.path.join(ROOT, "profiles", f"ncu_{tag}_summary.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return {"warp_inst": float(d["smsp__inst_executed.sum"]["value"]), "dram_bytes": float(d["dram_bytes_per_launch"]),
                "lanes_per_inst": float(d["smsp__thread_inst_executed_per_inst_executed.ratio"]["value"]),
                "kernel_ms_under_ncu": float(d["gpu__time_duration.sum"]["value"]), "source": os.path.relpath(p, ROOT),
                "kernel_version": d.get("kernel_version"), "note": d.get("note")}
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


# ----------------------------------------------------------------------------------------- oracle legs (checker only)
def oracle_render_timed(oracle, scene, spp, threads=0):
    t0 = time.perf_counter()
    r = oracle.render(scene, scene.integrator(), seed=0, spp=spp, n_threads=threads)
    return r, time.perf_counter() - t0


def oracle_rate(mitr, oracle, width, height, bins, spp, max_depth, threads=0):
    """Time one oracle render of the Cornell box; returns (M-paths/s, seconds, result)."""
    scene = mitr.load_dict(mitr.cornell_box_config(width, height, bins, spp, max_depth))
    r, dt = oracle_render_timed(oracle, scene, spp, threads)
    return width * height * spp / dt / 1e6, dt, r


def pick_cpu_sample(mitr, oracle, target_s):
    """Largest square sub-resolution of C2 (same spp, bins, depth) the oracle renders in ~target_s."""
    oracle_rate(mitr, oracle, 32, 32, C2["temporal_bins"], 16, C2["max_depth"])            # warm-up (C1-sized)
    rate, _, _ = oracle_rate(mitr, oracle, 32, 32, C2["temporal_bins"], C2["spp"], C2["max_depth"])
    paths = rate * 1e6 * target_s
    side = int((paths / C2["spp"]) ** 0.5) // 16 * 16
    return max(32, min(C2["width"], side))


def film_parity_device(torch, film_dev, ref_np, rows=32):
    """BASELINE's parity metric on the device: per-time-bin and whole-film relative L1 between the GPU film (CUDA tensor,
    (H,W,T,C)) and the oracle's film (numpy), streamed over row blocks.  Same definition as tests/parity_utils.film_l1."""
    H, T = film_dev.shape[0], film_dev.shape[2]
    num = torch.zeros(T, dtype=torch.float64, device="cuda")
    den = torch.zeros_like(num); mag = torch.zeros_like(num)
    for y0 in range(0, H, rows):
        r = torch.from_numpy(ref_np[y0:y0 + rows]).cuda().double()
        g = film_dev[y0:y0 + rows].double()
        num += (g - r).abs().sum(dim=(0, 1, 3)); den += r.abs().sum(dim=(0, 1, 3)); mag += g.abs().sum(dim=(0, 1, 3))
    nz = den > 0
    per_t = (num[nz] / den[nz])
    return {"whole": float(num.sum() / max(float(den.sum()), 1e-300)),
            "max_t": float(per_t.max()) if per_t.numel() else 0.0,
            "mean_t": float(per_t.mean()) if per_t.numel() else 0.0,
            "zero_bins_mismatch": int(((mag > 0) != nz).sum()), "bins_with_energy": int(nz.sum())}


def counters_dict(t):
    names = ("paths", "vertices", "deposits", "deposits_in_range", "bin_checksum", "shadow_rays")
    return {k: int(v) for k, v in zip(names, t)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import mitransient_b200 as mitr
    from oracle import oracle
    oracle.build()
    cores = len(os.sched_getaffinity(0))
    side = pick_cpu_sample(mitr, oracle, target_s=8.0)      # the full 512x512 job on a box that renders it in <= ~8 s per step
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
        "config": {"workload": "C2 cornell_box transient_path 512x512x1024 256spp depth8" + ("" if side == C2["width"] else " (bounded CPU sample)"),
                   "note": "mitsuba/drjit not installable here: CPU oracle port of the reference path, scalar C++"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------- roofline
def film_passes(integ, info=None):
    """DRAM passes over the film of the last launch: 3 on the global-reduction path behind a memset, else 1."""
    info = info or integ.last_render_info()
    volumetric = type(integ).__name__ == "TransientPRBVolpathIntegrator"      # zeroes its film inside the kernel (kZero)
    return 3 if info["histogram_path"] == "global" and not volumetric else 1


def roofline_block(n_v, W, H, T, kernel_ms, peak, peak_src, clocks, ncu, film_passes=1):
    """SURVEY 8(d) byte model next to what the kernel really does: its DRAM traffic is the film written once (path state
    lives in registers), its ceiling is instruction issue."""
    b_alg = 128.0 * n_v + 16.0 * W * H * T + 12.0 * W * H
    achieved = b_alg / (kernel_ms * 1e-3) / 1e9
    # DRAM traffic model: the shared-memory histogram path (and the volumetric kernel, which zeroes its own film in the L2)
    # writes the film once; the global-reduction path behind a memset writes it (memset), reads it back and writes it again
    film_bytes = film_passes * 16.0 * W * H * T + 12.0 * W * H
    r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
         "traffic": ncu["dram_bytes"] if ncu else None, "peak_source": peak_src,
         "kernel": "render_transient_path_kernel", "kernel_ms": kernel_ms, "algorithmic_bytes": b_alg,
         "live_path_vertices": n_v,
         "model": "SURVEY 8d: 128 B x N_v + 16 B x H*W*T + 12 B x H*W (wavefront byte model).  This kernel keeps path state "
                  "in registers: its real DRAM traffic is the film (dram_frac: written once by the shared-memory histogram path; memset + "
                  "read + write = 3 passes by the global-reduction path), its ceiling is instruction issue",
         "dram_frac": film_bytes / (kernel_ms * 1e-3) / 1e9 / peak,
         "dram_bytes_model": film_bytes, "film_passes": film_passes}
    if ncu:
        f_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        slots = N_SM * 4 * f_mhz * 1e6 * kernel_ms * 1e-3
        r["issue"] = {"frac": ncu["warp_inst"] / slots, "warp_inst": ncu["warp_inst"], "lanes_per_inst": ncu["lanes_per_inst"],
                      "sm_mhz": f_mhz, "issue_slots": slots,
                      "what": "executed warp instructions (ncu capture of the same workload, committed) / (148 SMs x 4 "
                              "schedulers x SM clock x kernel time of THIS run)", "source": ncu["source"],
                      "kernel_version": ncu.get("kernel_version"), "note": ncu.get("note")}
        r["traffic_source"] = ncu["source"] + " (ncu --set full of the same command; not re-measured in this run)"
    else:
        r["issue"] = None
    return r


def time_render(torch, lib, _lib, C, dev, params, film, steady, stream, steps, warmup):
    for _ in range(warmup):
        _lib.check(lib.tr_render(dev.handle, C.byref(params), film.data_ptr(), steady.data_ptr(), None, stream.cuda_stream))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        _lib.check(lib.tr_render(dev.handle, C.byref(params), film.data_ptr(), steady.data_ptr(), None, stream.cuda_stream))
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def extra_workload(torch, mitr, lib, _lib, C, oracle, name, tag, make_scene, W, H, T, spp, parity_spp, cpu_spp, steps, peak, peak_src,
                   clocks):
    """One of BASELINE.json's other single-GPU configurations: kernel time at the full size, exact counters, roofline
    block, the oracle's rate on a bounded sample (same film, fewer spp) and GPU-vs-oracle parity at the FULL film size."""
    stream = torch.cuda.current_stream()
    scene = mitr.load_dict(make_scene(spp))
    integ, sensor = scene.integrator(), scene.sensors()[0]
    dev = integ.device_scene(scene, sensor)
    film = torch.zeros((H, W, T, 4), dtype=torch.float32, device="cuda")
    steady = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
    counters = torch.zeros(8, dtype=torch.int64, device="cuda")
    params = integ.make_params(spp, spp, 0, film_channels=4)
    integ.fill_params_(params, scene, sensor)
    _lib.check(lib.tr_render(dev.handle, C.byref(params), film.data_ptr(), steady.data_ptr(), counters.data_ptr(), stream.cuda_stream))
    torch.cuda.synchronize()
    cnt = counters_dict(counters.cpu().tolist())
    ms = time_render(torch, lib, _lib, C, dev, params, film, steady, stream, steps, 3)
    out = {"workload": name, "paths_per_step": W * H * spp, "kernel_ms": ms, "value": W * H * spp / ms / 1e3, "unit": UNIT,
           "steps": steps, "counters": cnt, "kernel": integ.last_render_info(),
           "roofline": roofline_block(cnt["vertices"], W, H, T, ms, peak, peak_src, clocks, ncu_summary(tag), film_passes(integ))}
    if oracle is not None:
        # parity at the full film size with fewer samples per pixel (the oracle has to finish in seconds)
        sp = mitr.load_dict(make_scene(parity_spp))
        ip, sn = sp.integrator(), sp.sensors()[0]
        dp = ip.device_scene(sp, sn)
        pp = ip.make_params(parity_spp, parity_spp, 0, film_channels=4)
        ip.fill_params_(pp, sp, sn)
        counters.zero_()
        _lib.check(lib.tr_render(dp.handle, C.byref(pp), film.data_ptr(), steady.data_ptr(), counters.data_ptr(), stream.cuda_stream))
        torch.cuda.synchronize()
        got = counters_dict(counters.cpu().tolist())
        ref, dt = oracle_render_timed(oracle, sp, parity_spp)
        par = film_parity_device(torch, film, ref["transient"])
        par.update({"counters_equal": got == ref["counters"], "spp": parity_spp, "film": [H, W, T, 4],
                    "what": f"GPU vs CPU oracle, same seed, {W}x{H}x{T} film, {parity_spp} spp"})
        out["parity"] = par
        if cpu_spp == parity_spp:
            rate = W * H * parity_spp / dt / 1e6
        else:
            sc2 = mitr.load_dict(make_scene(cpu_spp))
            _, dt = oracle_render_timed(oracle, sc2, cpu_spp)
            rate = W * H * cpu_spp / dt / 1e6
        out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": len(os.sched_getaffinity(0)), "kind": "port",
                               "sample": f"{name}, {cpu_spp} spp ({W * H * cpu_spp} paths, {dt:.1f} s); CPU oracle restatement "
                                         "(scalar C++), NOT Dr.Jit llvm_ad_rgb"}
        del ref
    del film, steady
    return out


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
        except Exception as exc:                               # no P2P / IPC on this box: fall back to the reduce
            if rank == 0:
                print(f"peer film unavailable ({exc}); falling back to NCCL reduce", file=sys.stderr)
            mode, peer = "reduce", None
        flag = torch.tensor([1.0 if mode == "peer" else 0.0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if flag.item() == 0.0:
            mode, peer = "reduce", None
    own_film = torch.zeros((H, W, T, 4), dtype=torch.float32, device="cuda") if (peer is None or n > 1) else None
    own_steady = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
    counters = torch.zeros(8, dtype=torch.int64, device="cuda")
    launches = [0]

    def targets(m):
        return (peer.film, peer.steady) if m == "peer" else (own_film, own_steady)

    def finish_step(m):
        if m == "peer":
            peer.fence()
        elif m == "reduce":
            dist.reduce(own_film, dst=0)
            dist.reduce(own_steady, dst=0)

    def step(m, params, with_counters=False):
        film, steady = targets(m)
        if m == "reduce" and rank == 0:
            film.zero_()                                       # last step's reduce landed in rank 0's film
        _lib.check(lib.tr_render(dev.handle, C.byref(params), film.data_ptr(), steady.data_ptr(),
                                 counters.data_ptr() if with_counters else None, stream.cuda_stream))
        launches[0] += 1
        finish_step(m)

    def timed(m, params, steps, warmup):
        """K steps bracketed by barrier + synchronize, CUDA events on the launching stream; returns
        (ms per step incl. film assembly, kernel ms per step) as the max over ranks, and the per-rank kernel times."""
        film, steady = targets(m)
        for _ in range(warmup):
            step(m, params)
        if n > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        kern = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        ev[0].record(stream)
        for i in range(steps):
            if m == "reduce" and rank == 0:
                film.zero_()
            kern[i][0].record(stream)
            _lib.check(lib.tr_render(dev.handle, C.byref(params), film.data_ptr(), steady.data_ptr(), None, stream.cuda_stream))
            kern[i][1].record(stream)
            launches[0] += 1
            finish_step(m)
        ev[1].record(stream)
        torch.cuda.synchronize()
        if n > 1:
            dist.barrier()
        t = torch.tensor([ev[0].elapsed_time(ev[1]) / steps, sum(a.elapsed_time(b) for a, b in kern) / steps],
                         dtype=torch.float64, device="cuda")
        per_rank = [float(t[1])]
        if n > 1:
            g = [torch.zeros_like(t) for _ in range(n)]
            dist.all_gather(g, t)
            per_rank = [float(x[1]) for x in g]
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), per_rank

    # ---- exact counters (N_v) from one untimed step; the film of this step is what `parity` checks
    params = integ.make_params(spp, spp, 0, rank=rank, world=n, film_channels=4)
    step(mode, params, with_counters=True)
    torch.cuda.synchronize()
    cnt = counters.cpu().tolist()
    if n > 1:
        ct = counters.clone(); dist.all_reduce(ct); cnt = ct.cpu().tolist()
    gpu_counters = counters_dict(cnt)
    paths_total, n_v = cnt[0], cnt[1]
    assert paths_total == W * H * spp, (paths_total, W * H * spp)

    # ---- the headline: W warm-up steps (the counting step was one), K timed steps
    clocks = ClockSampler(local_rank)
    for _ in range(max(0, args.warmup - 1)):
        step(mode, params)
    if rank == 0:
        clocks.start()
    launches[0] = 0
    ms_per_step, kern_ms, kern_ms_per_rank = timed(mode, params, args.steps, 0)
    clk = clocks.stop() if rank == 0 else None
    headline_launches = launches[0]
    value = paths_total / (ms_per_step * 1e-3) / 1e6
    render_info = integ.last_render_info()          # which histogram path / CTA shape the timed launches used

    # ---- N > 1: the assembled film against a single-GPU render of the same job, then the other multi-GPU records
    parity = None
    multi = None
    if n > 1:
        if rank == 0:
            film_asm = torch.as_tensor(peer.film, device="cuda") if mode == "peer" else own_film
            ref_film = torch.zeros((H, W, T, 4), dtype=torch.float32, device="cuda")
            ref_steady = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
            c1 = torch.zeros(8, dtype=torch.int64, device="cuda")
            p1 = integ.make_params(spp, spp, 0, rank=0, world=1, film_channels=4)
            _lib.check(lib.tr_render(dev.handle, C.byref(p1), ref_film.data_ptr(), ref_steady.data_ptr(), c1.data_ptr(), stream.cuda_stream))
            torch.cuda.synchronize()
            num = (film_asm.double() - ref_film.double()).abs().sum(dim=(0, 1, 3)) if False else None   # (streamed below)
            Tn = torch.zeros(T, dtype=torch.float64, device="cuda"); Td = torch.zeros_like(Tn)
            for y0 in range(0, H, 32):
                a_, b_ = film_asm[y0:y0 + 32].double(), ref_film[y0:y0 + 32].double()
                Tn += (a_ - b_).abs().sum(dim=(0, 1, 3)); Td += b_.abs().sum(dim=(0, 1, 3))
            nz = Td > 0
            parity = {"what": f"film assembled from {n} ranks ({mode}) vs the same {spp}-spp job rendered by rank 0 alone "
                              "(identical RNG streams: global sample index)",
                      "counters_equal": counters_dict(c1.cpu().tolist()) == gpu_counters,
                      "whole": float(Tn.sum() / Td.sum()), "max_t": float((Tn[nz] / Td[nz]).max()),
                      "bit_identical": bool(torch.equal(film_asm, ref_film))}
            del ref_film, ref_steady
        dist.barrier()
        multi = {}
        ksteps = max(3, min(args.steps, 5))
        other = "reduce" if mode == "peer" else ("peer" if peer is not None else None)
        if other:
            o_ms, o_k, o_pr = timed(other, params, ksteps, 1)
            multi["weak_" + other] = {"ms_per_step": o_ms, "kernel_ms": o_k, "kernel_ms_per_rank": [round(x, 3) for x in o_pr],
                                      "value": paths_total / (o_ms * 1e-3) / 1e6, "assembly_ms": o_ms - o_k,
                                      "what": f"same weak-scaling step with mode={other}"}
        multi["weak_" + mode] = {"ms_per_step": ms_per_step, "kernel_ms": kern_ms, "assembly_ms": ms_per_step - kern_ms}
        # strong scaling: C2 fixed at 256 spp, N ranks.  Every rank but rank 0 pushes its (N-1)/N share of the 4.29 GB
        # film into rank 0 over NVLink while it renders for ~17/N ms: rank 0's NVLink ingest is the limiter.
        ps = integ.make_params(C2["spp"], C2["spp"], 0, rank=rank, world=n, film_channels=4)
        for m in ([mode] + ([other] if other else [])):
            s_ms, s_k, s_pr = timed(m, ps, ksteps, 1)
            ingest = (n - 1) / n * (16.0 * W * H * T) / (s_ms * 1e-3) / 1e9
            multi["strong_" + m] = {"spp": C2["spp"], "ms_per_step": s_ms, "kernel_ms_per_rank": [round(x, 3) for x in s_pr],
                                    "value": W * H * C2["spp"] / (s_ms * 1e-3) / 1e6, "rank0_ingest_GBps": ingest,
                                    "what": "C2 at 256 spp split over N ranks; ms_per_step includes the film assembly on rank 0"}
            if m == "reduce":
                # the same kernels without any assembly: every rank keeps its own tiles in its own HBM (the device-resident
                # default when the consumer is sharded too -- the layout tr_render_to_host with world > 1 copies out)
                multi["strong_sharded"] = {"spp": C2["spp"], "ms_per_step": s_k, "kernel_ms_per_rank": [round(x, 3) for x in s_pr],
                                           "value": W * H * C2["spp"] / (s_k * 1e-3) / 1e6,
                                           "what": "C2 at 256 spp split over N ranks, film left sharded (no assembly): the slowest "
                                                   "rank's kernel time of the reduce-mode steps"}
        if n == 8:                                             # BASELINE config 5 as written: 4096 spp over 8 GPUs
            p5 = integ.make_params(4096, 4096, 0, rank=rank, world=n, film_channels=4)
            for m in ([mode] + ([other] if other else [])):
                c_ms, c_k, c_pr = timed(m, p5, 3, 1)
                multi["C5_" + m] = {"spp": 4096, "ms_per_step": c_ms, "kernel_ms": c_k, "assembly_ms": c_ms - c_k,
                                    "kernel_ms_per_rank": [round(x, 3) for x in c_pr],
                                    "value": W * H * 4096 / (c_ms * 1e-3) / 1e6,
                                    "what": "BASELINE config 5: Cornell box 4096 spp, pixel tiles over 8 GPUs, film on rank 0"}

    # ---- e2e: host buffers; blob upload (H2D) + render + film to pinned host memory inside the timed region.
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
        blob_bytes = dev.info()["blob_bytes"]

        def e2e_step():
            _lib.check(lib.tr_scene_refresh(dev.handle, stream.cuda_stream))
            _lib.check(lib.tr_render_to_host(dev.handle, C.byref(p_e2e), host_T.data_ptr(), host_S.data_ptr(),
                                             None, stream.cuda_stream))
        for _ in range(2):
            e2e_step()
        k_e2e = max(2, min(args.steps, 5))
        if n > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            e2e_step()                                         # synchronous: returns when this rank's shard is on the host
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if n > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)          # a step ends when the slowest rank's shard is on the host
        dt = te.item()
        nb = torch.tensor([float(host_T.numel() * 4 + host_S.numel() * 4)], dtype=torch.float64, device="cuda")
        if n > 1:
            dist.all_reduce(nb)
        e2e = {"value": W * H * spp * k_e2e / dt / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(blob_bytes * n),
               "d2h_bytes_per_step": int(nb.item()),
               "steps": k_e2e, "ms_per_step": 1e3 * dt / k_e2e,
               "what": "tr_scene_refresh (scene blob H2D) + tr_render_to_host: developed (H,W,T,3) film + steady into pinned host "
                       "memory, D2H overlapped with rendering band by band" +
                       (f"; the film is sharded by pixel tiles: each of the {n} ranks renders its tiles and copies its "
                        "share to its own pinned host buffer over its own PCIe link (bytes summed over ranks; time = the "
                        "slowest rank's K steps, no per-step barrier)" if n > 1 else "")}
        del host_T, host_S
    except Exception as exc:  # keep the headline line even if the host path fails
        e2e = {"value": None, "unit": UNIT, "error": str(exc)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant (only) kernel: render_transient_path_kernel
    peak, peak_src = measured_peak_hbm()
    roofline = roofline_block(n_v / n, W, H / n, T, kern_ms, peak, peak_src, clk, ncu_summary("C2") if n == 1 else None,
                              film_passes(integ, render_info))
    roofline["kernel_ms_per_rank"] = [round(x, 3) for x in kern_ms_per_rank]

    # ---- CPU baseline (oracle port), rank 0, N = 1 only: the FULL C2 render when the host finishes it in ~20 s -- then it
    # doubles as the parity check at the quoted size -- else a bounded sub-resolution (rate only)
    cpu, extra = None, None
    if n == 1 and not args.no_cpu_baseline:
        try:
            from oracle import oracle
            oracle.build()
            cores = len(os.sched_getaffinity(0))
            side = pick_cpu_sample(mitr, oracle, target_s=20.0)
            rate, dt, ref = oracle_rate(mitr, oracle, side, side, T, C2["spp"], C2["max_depth"])
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"Cornell box {side}x{side} px x {T} bins x {C2['spp']} spp, max_depth 8 "
                             f"({side * side * C2['spp']} paths, {dt:.1f} s); CPU oracle restatement (scalar C++), "
                             "NOT Dr.Jit llvm_ad_rgb (mitsuba is not installable here)"}
            if side == W:
                par = film_parity_device(torch, own_film, ref["transient"])
                par.update({"counters_equal": gpu_counters == ref["counters"], "gpu_counters": gpu_counters,
                            "oracle_counters": ref["counters"], "tolerance": {"whole": 1e-5, "max_t": 1e-4},
                            "what": "GPU film of the counting step vs the CPU oracle's render of the same C2 job (same seed): "
                                    "exact integer work counters, per-time-bin and whole-film relative L1"})
                par["ok"] = bool(par["counters_equal"] and par["whole"] <= 1e-5 and par["max_t"] <= 1e-4
                                 and par["zero_bins_mismatch"] == 0)
                parity = par
            else:
                parity = {"skipped": f"the oracle renders only {side}x{side} of C2 in 20 s on {cores} cores; parity at the "
                                     "quoted size is covered by tests/test_gpu_fullsize.py"}
            del ref
            if not args.no_extra:
                del own_film
                torch.cuda.empty_cache()
                extra = []
                xs = max(3, min(args.steps, 5))
                extra.append(extra_workload(
                    torch, mitr, lib, _lib, C, oracle, "C3 nlos_z transient_nlos_path confocal 256x256x4096, 1024 spp, laser + "
                    "hidden-geometry sampling", "C3", lambda s: mitr.nlos_z_scene(256, 256, 4096, s), 256, 256, 4096, 1024, 16, 16,
                    xs, peak, peak_src, clk))
                extra.append(extra_workload(
                    torch, mitr, lib, _lib, C, oracle, "C4i cornell_box_volumetric transient_prbvolpath 256x384x512, 256 spp, "
                    "max_depth 16", "C4i", lambda s: mitr.cornell_box_volumetric(256, 384, 512, s), 256, 384, 512, 256, 16, 16,
                    xs, peak, peak_src, clk))
                extra.append(extra_workload(
                    torch, mitr, lib, _lib, C, oracle, "C2-mesh cornell box as 36 mesh triangles (the flavour of cbox_diffuse.xml: 18 "
                    "parallelograms, lock-step loop) transient_path 512x512x1024, 256 spp", "C2mesh",
                    lambda s: mitr.cornell_box_mesh_config(512, 512, 1024, s), 512, 512, 1024, 256, 4, 4, xs, peak, peak_src, clk))
                extra.append(extra_workload(
                    torch, mitr, lib, _lib, C, oracle, "C2-bvh cornell box + 1280-triangle icosphere (1287 primitives, SAH BVH path) "
                    "transient_path 512x512x1024, 256 spp", "C2bvh",
                    lambda s: mitr.cornell_box_sphere_config(512, 512, 1024, s, 3), 512, 512, 1024, 256, 1, 1, 3, peak, peak_src, clk))
        except Exception as exc:
            import traceback
            traceback.print_exc(file=sys.stderr)
            cpu = cpu or {"value": None, "unit": UNIT, "error": str(exc)}

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
                   "kernel": dict(render_info, note="histogram_path global = red.global.add.v4.f32 into the film, which a "
                                  "cudaMemsetAsync inside the timed step zeroes (counted in ms_per_step and roofline.kernel_ms, "
                                  "not in gpu_launches); shared = shared-memory slots flushed once per pixel"),
                   "l2": "no flush needed: each step writes the 4 GiB film (>> 126 MB L2); the scene lives in shared memory",
                   "parallelism": f"pixel-tile x{n}"},
        "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "e2e": e2e, "gpu_launches": headline_launches,
        "clocks": clk,
    }
    if multi is not None:
        line["multi_gpu"] = multi
    if extra is not None:
        line["extra"] = extra
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C3 / C4i / mesh workloads (N = 1)")
    ap.add_argument("--multi-gpu-mode", default="peer", choices=["peer", "reduce"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
