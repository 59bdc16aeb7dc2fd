#!/usr/bin/env python
"""bench.py -- the hot path of Crixus (BASELINE.json north_star) on N B200s of one node.

    python bench.py --gpus 1 --steps K --warmup W                 (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W  (the reference's algorithm on the host cores)

A step = one pass of the hot path (swap_normals, set_bound_elem, cell binning, calc_trisize, calc_vert_volume,
geometry fill) over the workload of BASELINE.json configs[1]: the SPHERIC-2 tank midpoint-subdivided x64 (~3.1e6
triangles) at dr/4, fluid filled through the shipped free-surface shape.  Metric: end-to-end preprocess throughput in
triangles/s (whole job); stage rates (vert-volume verts/s, fill cells/s, seconds) are reported beside it.

`value`  : inputs (the welded mesh) already resident in HBM when the timed region starts, CUDA-event timed.
`e2e`    : the same pass through the C-ABI call cx_preprocess_host on pinned HOST buffers: H2D of the raw STL content
           (corners + facet normals), bounding box + vertex weld on the device, D2H of every result array -- all
           inside the timed region.
`roofline`: calc_vert_volume is FP32-ALU bound: achieved = 68921*T*56 flop/vertex * verts/s against the
           FFMA peak measured in this run; the fill is reported against the measured HBM copy bandwidth.
`cpu_baseline`: the CPU oracle (oracle/liboracle.so, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (profiles/r01h_*.txt),
# valid for the default workload (config 2) on one GPU only
NCU_TRAFFIC = {"k_vert_volume": 135.112960e6 + 9.579264e6, "k_resolve_tiles": 140.846592e6 + 7.789312e6}

VOXELS = 41 ** 3
FLOP_PER_VOXEL_TRI = 56      # 25 cube + 8 voronoi + 15 wall + 8 edge


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_pipeline_rate(dr_sample, verbose=False):
    """Runs the oracle (all host threads) on a bounded sample of the workload; returns dict(value, sample, seconds...)."""
    from crixus_b200 import workloads as wl
    from oracle import orc
    O = orc.Oracle()
    w = wl.config2_sample(dr_sample)
    t0 = time.perf_counter()
    out = orc.run_pipeline(O, w["tris"], w["normals"], w["dr"], swap_normals=w["swap_normals"], periodic=w["periodic"],
                           zero_open=w["zero_open"], container=w["container"], fills=w["fills"], fshape=w["fshape"])
    dt = time.perf_counter() - t0
    return dict(value=w["nbe"] / dt, seconds=dt, cores=O.num_threads(), sample=w["name"], nbe=w["nbe"], nvert=w["nvert"],
                nfluid=int(out["nfluid"]))


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's algorithm (oracle port, the reference has no CPU path of its own) on the host
    cores.  Rank 0 alone works; each step is a bounded sample of config 2 sized from --steps/--warmup."""
    if rank != 0:
        return
    budget = 150.0 / max(args.steps + args.warmup, 1)          # seconds per step
    # the oracle runs ~1.5e3 triangles/s on 8 threads at h=0.625dr; tank area ~11.7 m^2 -> tris = 2*11.7/(0.625dr)^2
    tris_target = max(1500.0 * budget * 0.8, 3000.0)
    dr_s = float(np.sqrt(2 * 11.7 / tris_target) / 0.625)
    dr_s = min(max(dr_s, 0.02), 0.12)
    rates = []
    for i in range(args.warmup + args.steps):
        r = cpu_pipeline_rate(dr_s)
        if i >= args.warmup:
            rates.append(r)
    tot_t = sum(r["seconds"] for r in rates)
    val = sum(r["nbe"] for r in rates) / tot_t
    line = {"impl": "reference", "metric": "end-to-end preprocess triangles/s (hot path)", "value": val, "unit": "triangles/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / len(rates),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1] (spheric2 x64 at dr/4, fshape fill); timed on a bounded sample: " + rates[0]["sample"]},
            "cpu_baseline": {"value": val, "unit": "triangles/s", "cores": rates[0]["cores"], "kind": "port", "sample": rates[0]["sample"]},
            "e2e": {"value": val, "unit": "triangles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--levels", type=int, default=3, help="midpoint subdivisions of the SPHERIC-2 base mesh (3 = config 2)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-dr", type=float, default=0.045, help="dr of the CPU-baseline sample")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from crixus_b200 import api, workloads as wl
    from crixus_b200.dist import ShardedHotPath

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (crixus_b200 has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = api.Context(local)
    dev = ctx.device
    w = wl.config2(args.levels)
    hp = ShardedHotPath(ctx, w, rank, world)
    hbm_peak, hbm_src = peaks()
    fp32_peak = ctx.measure_fp32_peak(5)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value)
    for _ in range(args.warmup):
        hp.step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ctx.launches()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    stage_acc = np.zeros(len(hp.STAGES))
    barrier()
    ev[0].record()
    for _ in range(args.steps):
        stage_acc += hp.step(timed=True)
    ev[1].record()
    barrier()
    ms = ev[0].elapsed_time(ev[1])
    launches = ctx.launches() - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    stage_ms = hp.collect_stage_ms() / args.steps          # max over ranks inside
    hp.vv_kernel_ms = float(hp._stage_rank_local[hp.STAGES.index("vert_volume")]) / args.steps
    nbe, nvert = w["nbe"], hp.nalive
    G = int(np.prod(hp.dimg))
    value = nbe * args.steps / (ms * 1e-3)

    # ---- end to end through cx_preprocess_host on pinned host buffers (e2e), rank-local full pass at N=1;
    #      at N>1 every rank runs its shard through the sharded driver with host<->device copies in the timed region
    e2e = hp.e2e(args.steps, barrier)

    line = None
    if rank == 0:
        T = hp.mean_trisize
        vv_ms = stage_ms[hp.STAGES.index("vert_volume")]
        fill_ms = stage_ms[hp.STAGES.index("fill")]
        flops_per_launch = VOXELS * T * FLOP_PER_VOXEL_TRI * hp.local_verts      # rank 0's launch
        achieved = flops_per_launch / (hp.vv_kernel_ms * 1e-3) / 1e12
        fill_bytes = 0.25 * G + 8.0 * hp.nfluid + 48.0 * nbe + 16.0 * w["nvert"] + 4.0 * hp.ncell3
        line = {
            "metric": "end-to-end preprocess triangles/s (hot path)", "value": value, "unit": "triangles/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "nbe": nbe, "nvert": w["nvert"], "lattice": [int(x) for x in hp.dimg],
                       "nfluid": hp.nfluid, "l2": "inputs+outputs of every stage > 126 MB L2 (sorted geometry 0.17 GB); "
                       "the fill's 1.2 MB bitfield is L2-resident by nature", "parallelism": hp.parallelism},
            "stages": {"preprocess_s": ms / args.steps * 1e-3,
                       "vert_volume_verts_per_s": nvert / (vv_ms * 1e-3), "fill_cells_per_s": G / (fill_ms * 1e-3),
                       "ms": {k: float(v) for k, v in zip(hp.STAGES, stage_ms)}, "fill_rounds": hp.fill_rounds},
            "roofline": {"kernel": "k_vert_volume", "bound": "fp32_alu", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp32_peak, "traffic": NCU_TRAFFIC["k_vert_volume"] if (world == 1 and args.levels == 3) else None,
                         "frac_note": "ALGORITHMIC flops of the reference's brute-force 41^3 quadrature over the measured FFMA peak; the kernel "
                                      "removes > 90 % of them by exact culling (monotone threshold search), hence frac > 1.  Executed view (ncu, "
                                      "profiles/r01h_vv.txt): 70e3 warp instructions per vertex, 2.7 instructions per cycle and SM, DRAM traffic 0.15 GB "
                                      "per launch against 23 GB algorithmic bytes (L2 hit 99.9 %)",
                         "peak_source": "FFMA micro-kernel measured in this run (cx_measure_fp32_peak); nominal 148*128*2*1.965 GHz = 74.4",
                         "flops_per_vertex": VOXELS * T * FLOP_PER_VOXEL_TRI, "mean_trisize": T,
                         "hbm_frac": (hp.vv_bytes / (hp.vv_kernel_ms * 1e-3) / 1e9) / hbm_peak},
            "roofline_fill": {"kernel": "k_resolve_tiles + k_fill_tiles rounds (whole geometry fill)", "bound": "hbm", "achieved": fill_bytes / (fill_ms * 1e-3) / 1e9,
                              "peak": hbm_peak, "unit": "GB/s", "frac": fill_bytes / (fill_ms * 1e-3) / 1e9 / hbm_peak,
                              "traffic": NCU_TRAFFIC["k_resolve_tiles"] if (world == 1 and args.levels == 3) else None,
                              "peak_source": hbm_src, "bytes": fill_bytes,
                              "frac_note": "the fill is bound by the directed tests of the cells near walls (k_resolve_tiles, ALU/latency), "
                                           "not by HBM; profiles/r01h_resolve.txt"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if not args.no_cpu and world == 1:
            r = cpu_pipeline_rate(args.cpu_dr)
            line["cpu_baseline"] = {"value": r["value"], "unit": "triangles/s", "cores": r["cores"], "kind": "port",
                                    "sample": r["sample"] + " (%.1f s)" % r["seconds"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
