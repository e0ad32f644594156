#!/usr/bin/env python
"""bench.py — Mrays/s closest-hit on BASELINE.json config 2 (+ BVH build ms), vs a CPU baseline.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on host cores

Workload (config.workload = "config2"): icosphere subdivision 7 (327 680 triangles),
3840x2160 pinhole camera rays (8 294 400 rays, origin a stride-0 broadcast as in the
reference's test/performance_test.py:36-41), `intersects_closest` incl. location + uv.
One step = one pass of the hot path over that ray batch.  N > 1: weak scaling — the BVH is
built on rank 0 and NCCL-broadcast, every rank traces its own full frame (camera shifted per
rank), no data-path collective.

Timing: W >= 3 warm-up steps; every timed step is bracketed by CUDA events on the launch
stream with an L2 flush (512 MiB memset) between steps; the job time is the max over ranks of
the summed step times.  Clocks are sampled through NVML during the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "trimesh-ray-optix_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

METRIC = "Mrays/s closest-hit"
UNIT = "Mrays/s"
WIDTH, HEIGHT = 3840, 2160
SUBDIV = 7


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def make_workload(device, rank: int):
    from triro import synth

    v, f = synth.icosphere(SUBDIV)
    # rank-dependent camera so that ranks do not trace identical rays
    o, d = synth.pinhole_rays(WIDTH, HEIGHT, device=device, origin=(0.02 * rank, -0.01 * rank, 3.0))
    return v, f, o, d


def bytes_per_ray(stats: dict, broadcast_origin: bool) -> float:
    b_in = 12.0 if broadcast_origin else 24.0
    return b_in + 26.0 + 80.0 * stats["nodes_per_ray"] + 48.0 * stats["tris_per_ray"]


# ------------------------------------------------------------------------------------- CPU arm
def cpu_closest(v, f, o_np, d_np, sample: int, reps: int = 1):
    """Times the oracle port (binary32 mirror behind its binned-SAH BVH2, OpenMP over rays) on a
    bounded sample of the workload's rays.  Returns (Mrays/s, cores, sample description, seconds)."""
    from oracle import oracle

    n = len(d_np)
    stride = max(1, n // sample)
    d_s = np.ascontiguousarray(d_np[::stride])
    o_s = np.ascontiguousarray(np.broadcast_to(o_np, d_np.shape)[::stride])
    mesh = oracle.OracleMesh(v, f, use_bvh=True)           # BVH build is not timed (neither is the GPU's)
    oracle.query(mesh, o_s[:1000], d_s[:1000], oracle.MIRROR, closest_only=True, want=("hit", "tri", "loc", "uv", "front"))
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        oracle.query(mesh, o_s, d_s, oracle.MIRROR, closest_only=True, want=("hit", "front", "tri", "loc", "uv"))
        best = min(best, time.perf_counter() - t0)
    return len(d_s) / best / 1e6, oracle.num_threads(), f"every {stride}th ray of the {WIDTH}x{HEIGHT} frame ({len(d_s)} rays)", best


def run_reference(args):
    """--impl reference: the reference's own CPU path is trimesh+embree (absent offline) and its GPU path
    needs the OptiX SDK (absent), so this arm times the oracle port of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from triro import synth

    v, f = synth.icosphere(SUBDIV)
    o, d = synth.pinhole_rays(WIDTH, HEIGHT, device="cpu")
    d_np = d.reshape(-1, 3).numpy()
    o_np = np.array([[0.0, 0.0, 3.0]], np.float32)
    from oracle import oracle

    mesh = oracle.OracleMesh(v, f, use_bvh=True)
    sample = 1_000_000
    stride = max(1, len(d_np) // sample)
    d_s = np.ascontiguousarray(d_np[::stride]); o_s = np.ascontiguousarray(np.broadcast_to(o_np, d_np.shape)[::stride])
    want = ("hit", "front", "tri", "loc", "uv")
    for _ in range(max(args.warmup, 1)):
        oracle.query(mesh, o_s, d_s, oracle.MIRROR, closest_only=True, want=want)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.query(mesh, o_s, d_s, oracle.MIRROR, closest_only=True, want=want)
    dt = (time.perf_counter() - t0) / args.steps
    val = len(d_s) / dt / 1e6
    cores = oracle.num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config2: icosphere subdiv 7 (327680 tris), 3840x2160 pinhole rays, intersects_closest",
                   "step": f"bounded sample: every {stride}th ray ({len(d_s)} rays) per step"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"every {stride}th ray of the frame ({len(d_s)} rays), oracle binary32 mirror + binned-SAH BVH2, OpenMP"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference OptiX path not buildable offline (no OptiX SDK); trimesh/embree absent; CPU arm = oracle port",
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch.distributed as dist

    from triro.backend import ops as hops
    from triro.ray.ray_optix import RayMeshIntersector

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # one process per GPU: keep this rank's pinned buffers and copy submission on the GPU's own NUMA node
    all_cpus = os.sched_getaffinity(0)
    from triro.distributed import bind_to_device_cpus

    local_cpus = None if os.environ.get("TRIRO_BENCH_NO_BIND") else bind_to_device_cpus(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    v, f, o, d = make_workload(dev, rank)
    vt, ft = torch.from_numpy(v), torch.from_numpy(f)
    bcast_ms = None
    if world > 1:
        from triro.distributed import ShardedRayMeshIntersector

        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        sharded = ShardedRayMeshIntersector.build(vt, ft, src=0)
        torch.cuda.synchronize(); dist.barrier()
        bcast_ms = (time.perf_counter() - t0) * 1e3
        rmi = sharded.local
    else:
        rmi = RayMeshIntersector(vertices=vt, faces=ft)
    # BVH build time (device time of the build pipeline alone), measured on every rank's GPU
    accel = hops.AccelStructure()
    vd, fd = vt.to(dev), ft.to(dev)
    build_ms = []
    for _ in range(4):
        accel.build(vd, fd, timing=True)
        build_ms.append(accel.build_ms)
    build_ms = min(build_ms[1:])
    accel.free()

    n = d.numel() // 3
    stats = hops.trace_stats(rmi.as_wrapper, o, d, "closest")
    bpr = bytes_per_ray(stats, broadcast_origin=True)

    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    steps, warm = args.steps, max(args.warmup, 3)

    def one_step():
        return rmi.intersects_closest(o, d)

    for _ in range(warm):
        res = one_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with ClockSampler(local_rank) as clk:
        for a, b in ev:
            flush.zero_()                       # evict the previous step's working set from L2 (untimed)
            a.record()
            res = one_step()
            b.record()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(step_ms)

    # end-to-end through the host-buffer entry point: pinned host rays in, pinned host results out
    o_host = torch.tensor([0.02 * rank, -0.01 * rank, 3.0]).pin_memory()
    d_host = d.reshape(-1, 3).cpu().pin_memory()
    out = hops.host_closest(rmi.as_wrapper, o_host, d_host)
    for _ in range(2):
        out = hops.host_closest(rmi.as_wrapper, o_host, d_host, out=out)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = hops.host_closest(rmi.as_wrapper, o_host, d_host, out=out)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    # sanity: the host path returns the same answer as the device path
    assert int(out["hit"].sum()) == int(res[0].sum()), "host path and device path disagree"

    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        peak, peak_src = measured_peaks()
        ms_per_step = total_ms / steps
        value = world * n * steps / (total_ms * 1e-3) / 1e6
        kernel_ms = statistics.mean(step_ms)
        achieved = bpr * n / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                traffic = json.load(fh).get("k_trace_closest_config2_dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config2: icosphere subdiv 7 (327680 tris), 3840x2160 pinhole rays, intersects_closest (hit, front, tri, loc, uv)",
                       "rays_per_gpu": n, "l2": "flushed between steps (512 MiB memset); step working set 315 MB > 126 MB L2",
                       "parallelism": f"ray-sharded x{world}, BVH built on rank 0 and NCCL-broadcast" if world > 1 else "single GPU"},
            "bvh_build_ms": build_ms, "bvh_nodes": rmi.as_wrapper.header["n_nodes"], "bvh_depth": rmi.as_wrapper.header["depth"],
            "bvh_blob_mb": rmi.as_wrapper.header["used_bytes"] / 1e6,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "k_trace<closest>",
                         "bytes_per_ray": bpr, "nodes_per_ray": stats["nodes_per_ray"], "tris_per_ray": stats["tris_per_ray"],
                         "hit_fraction": stats["hit_fraction"], "kernel_ms": kernel_ms,
                         "roofline_mrays_per_s": peak * 1e9 / bpr / 1e6,
                         "compulsory_bytes_per_ray": 12.0 + 26.0 + rmi.as_wrapper.header["used_bytes"] / n},
            "e2e": {"value": world * n * steps / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": 12 * n + 12,
                    "d2h_bytes_per_step": 26 * n, "api": "rt_host_trace_closest (pinned host buffers, ramped chunks up to 2 Mi rays on 3 streams)"},
            "gpu_launches": steps * world, "clocks": clk.summary(),
        }
        if bcast_ms is not None:
            line["bvh_build_plus_broadcast_ms"] = bcast_ms
        line["host_cpus_bound"] = len(local_cpus) if local_cpus else None
        if world == 1:
            os.sched_setaffinity(0, all_cpus)      # the CPU baseline uses every host core
            val, cores, sample, secs = cpu_closest(v, f, np.array([[0.0, 0.0, 3.0]], np.float32), d.reshape(-1, 3).cpu().numpy(),
                                                   sample=2_000_000)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                    "seconds": secs}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
