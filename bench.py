#!/usr/bin/env python
"""bench.py -- the hot path of Crixus (BASELINE.json north_star) on N B200s of one node.

    python bench.py --gpus 1 --steps K --warmup W [--config C]      (N>1: launched by torch.distributed.run, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W  (the reference's own kernels on the host cores)

A step = one pass of the hot path (swap_normals, set_bound_elem, cell binning, periodic links, calc_trisize, calc_vert_volume,
fills) over the workload of BASELINE.json configs[C-1].  Default C = 4 = configs[3], the configuration BASELINE names for the
1/2/4/8-GPU runs: sphere in a tank, 48.3 M triangles, 1001^3-cell fluid lattice.  (--config 2 = the SPHERIC-2 x64 workload of
round 1, --config 3 the periodic channel, --config 5 the 198 M-triangle tank with a 1e10-cell lattice.)
Metric: end-to-end preprocess throughput in triangles/s (whole job); stage rates (vert-volume verts/s, fill cells/s, seconds)
are reported beside it.

`value`   : inputs (the welded mesh) already resident in HBM when the timed region starts, CUDA-event timed, max over ranks.
`e2e`     : the same pass through the C-ABI call on pinned HOST buffers holding the raw STL content (corners + facet normals):
            cx_preprocess_host at N=1, cx_preprocess_dist at N>1 (each rank uploads 1/N, NCCL all-gather, each rank downloads
            its 1/N share of the results) -- H2D, bounding box + weld on the device, all stages, D2H inside the timed region.
`check`   : popcount and order-independent 64-bit checksums of the fluid bitfield, the volume bits and the sorted geometry; at
            N>1 rank 0 repeats the pass through the single-GPU calls on the same inputs and the checksums must be identical.
`roofline`: calc_vert_volume is FP32-ALU bound: achieved = 68921*T*56 flop/vertex * verts/s against the FFMA
            peak measured in this run (plus the executed-instruction view from the committed ncu capture); `roofline_fill`
            reports the fill against the measured HBM copy bandwidth.
`cpu_baseline` / `--impl reference`: the reference's own src/crixus_d.cu compiled for the host (oracle/_ref, all host threads;
            the C port oracle/liboracle.so where the reference build is absent) on a bounded sample of the same workload --
            the same sample in both arms.
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
sys.path.insert(0, ROOT)

# per-launch figures of the committed `ncu --set full` captures (profiles/), valid for the named workload on one GPU:
# dram bytes (read + write), warp instructions executed, issue-slot utilisation, FP32-pipe share of the executed instructions
NCU = {}
try:
    NCU = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
except Exception:
    pass

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


# ------------------------------------------------------------------------------------------------ CPU arms
def host_threads():
    """all host cores, whatever OMP_NUM_THREADS the launcher exported (torch.distributed.run sets it to 1)"""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_backend():
    """the reference's own kernels compiled for the host (oracle/_ref/libcrixus_refshim.so) when that build exists, else the
    C restatement; both OpenMP over all host cores"""
    n = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(n)          # before the OpenMP runtimes of the oracle libraries start
    from oracle import orc
    if orc.RefShim.available():
        return orc.RefShim(threads=n), "reference", n
    O = orc.Oracle()
    return O, "port", O.num_threads()


def cpu_pass(B, sample):
    """one pass of the hot path over the sample on backend B -> seconds, nfluid"""
    from oracle import orc
    name, o, corners, normals = sample
    t0 = time.perf_counter()
    out = orc.run_pipeline(B, corners, normals, o["dr"], swap_normals=o["swap_normals"], periodic=o["periodic"], zero_open=o["zero_open"],
                           container=o["container"], fills=o["fills"], fshape=o["fshape"])
    return time.perf_counter() - t0, int(out["nfluid"]), int(out["nvert"])


def sample_text(sample, nvert=None):
    name, o, corners, normals = sample
    return "%s: %d triangles%s" % (name, corners.shape[0], "" if nvert is None else ", %d vertices" % nvert)


def run_reference_arm(args, rank):
    """--impl reference: the reference's kernels (src/crixus_d.cu compiled for the host, OpenMP) on all host cores.  Rank 0 alone
    works; each step is one pass over the bounded sample of the configured workload (the sample of the GPU arm's cpu_baseline).
    Nothing of crixus_b200's library is loaded here."""
    if rank != 0:
        return
    from crixus_b200 import workloads as wl     # generators only (numpy); the welded mesh comes from the backend's own weld
    B, kind, cores = cpu_backend()
    sample = wl.cpu_sample(args.config, args.steps + args.warmup)
    secs, nfl, nv = [], 0, 0
    for i in range(args.warmup + args.steps):
        dt, nfl, nv = cpu_pass(B, sample)
        if i >= args.warmup:
            secs.append(dt)
    nbe = sample[2].shape[0]
    val = nbe * len(secs) / sum(secs)
    name = wl.bench_config_name(args.config)
    line = {"impl": "reference", "metric": "end-to-end preprocess triangles/s (hot path)", "value": val, "unit": "triangles/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(secs) / len(secs),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "timed_on": "bounded sample -- " + sample_text(sample, nv), "nfluid": nfl},
            "cpu_baseline": {"value": val, "unit": "triangles/s", "cores": cores, "kind": kind, "sample": sample_text(sample, nv)},
            "e2e": {"value": val, "unit": "triangles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def reference_cuda_leg():
    """The reference's own CUDA build (oracle/_ref/Crixus_ref, sm_100) and our CLI on the same files: 763 776-triangle SPHERIC-2
    level (the largest the reference survives; it crashes at configs[1]).  Wall clock of the whole process, both pay the same
    CUDA start-up."""
    import tempfile
    ref = os.path.join(ROOT, "oracle", "_ref", "Crixus_ref")
    cli = os.path.join(ROOT, "crixus_b200", "bin", "crixus")
    if not (os.path.exists(ref) and os.path.exists(cli)):
        return {"unavailable": "oracle/_ref/Crixus_ref or crixus_b200/bin/crixus not built"}
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import time_reference_cuda as trc
        with tempfile.TemporaryDirectory() as tmp:
            return trc.compare_level(tmp, 2, parity=False)
    except Exception as e:     # reported, never fatal for the bench line
        return {"unavailable": "%s: %s" % (type(e).__name__, e)}


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", type=int, default=4, help="1-based index into BASELINE.json configs (default 4 = configs[3])")
    ap.add_argument("--size", type=int, default=None, help="override the size parameter of the workload family (debugging)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-check", action="store_true", help="skip the single-GPU re-run that the N>1 checksums are compared with")
    ap.add_argument("--no-reference-cuda", action="store_true", help="skip the reference's CUDA binary vs our CLI on the same files")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from crixus_b200 import api, workloads as wl
    from crixus_b200.dist import ShardedHotPath

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (crixus_b200 has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = api.Context(local)
    if world > 1:   # the library's own NCCL communicator: rank 0 creates the id, torch.distributed only ships the 128 bytes
        box = [api.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx.dist_init(box[0], world, rank)

    # ---- workload: rank 0 generates the raw STL content, the other ranks receive it; every rank keeps it in pinned host memory
    # (the input of the e2e call) and welds it on its own device (the device-resident input of `value`)
    t_setup = time.time()
    name, opts, gen = wl.bench_config(args.config, args.size)
    if rank == 0:
        corners, normals = gen()
        nbe = corners.shape[0]
    hdr = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == 0:
        hdr[0] = nbe
    if world > 1:
        dist.broadcast(hdr, 0)
    nbe = int(hdr.item())
    dcorn = torch.empty((nbe * 3, 3), dtype=torch.float32, device=dev)
    dnorm3 = torch.empty((nbe, 3), dtype=torch.float32, device=dev)
    if rank == 0:
        dcorn.copy_(torch.from_numpy(corners.reshape(-1, 3)))
        dnorm3.copy_(torch.from_numpy(normals))
        del corners, normals
    if world > 1:
        dist.broadcast(dcorn, 0)
        dist.broadcast(dnorm3, 0)
    hcorn = torch.empty((nbe * 3, 3), dtype=torch.float32).pin_memory()
    hnorm3 = torch.empty((nbe, 3), dtype=torch.float32).pin_memory()
    hcorn.copy_(dcorn); hnorm3.copy_(dnorm3)
    torch.cuda.synchronize()
    pos4, ep4, bbmin, bbmax = ctx.weld_device(dcorn, opts["dr"])
    nvert = int(pos4.shape[0])
    norm4 = torch.zeros((nbe, 4), dtype=torch.float32, device=dev)
    norm4[:, :3] = dnorm3
    w = dict(opts, name=name, nbe=nbe, nvert=nvert, pos=pos4.clone(), ep=ep4, norm=norm4, bbmin=bbmin, bbmax=bbmax)
    del dcorn, dnorm3, pos4
    torch.cuda.empty_cache()
    hp = ShardedHotPath(ctx, w, rank, world, byte_stats=(nbe <= 4_000_000))
    t_setup = time.time() - t_setup
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
    barrier()
    ev[0].record()
    for _ in range(args.steps):
        hp.step(timed=True)
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
    nalive = hp.nalive
    G = int(np.prod(hp.dimg))
    value = nbe * args.steps / (ms * 1e-3)

    # ---- bitwise evidence: checksums of this run; at N>1 rank 0 repeats the pass through the single-GPU calls
    check = hp.check()
    if world > 1:
        parts = [int(check["fpos_hash"], 16), int(check["vol_hash"], 16), int(check["geometry_hash"], 16), check["nfluid_popcount"]]
        halves = torch.tensor([x >> 32 for x in parts] + [x & 0xFFFFFFFF for x in parts], dtype=torch.int64, device=dev)
        lo, hi = halves.clone(), halves.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        check["identical_on_all_ranks"] = bool(torch.equal(lo, hi))
        if not args.no_check:
            if rank == 0:
                ref = hp.single_gpu_check()
                check["single_gpu_pass"] = ref
                check["equals_single_gpu_pass"] = all(ref[k] == check[k] for k in ref)
            barrier()
    # ---- end to end through the C-ABI on pinned host buffers (e2e)
    e2e = hp.e2e(args.steps, barrier, hcorn, hnorm3)

    if rank == 0:
        T = hp.mean_trisize
        vv_ms = stage_ms[hp.STAGES.index("vert_volume")]
        fill_ms = stage_ms[hp.STAGES.index("fill")]
        flops_per_launch = VOXELS * T * FLOP_PER_VOXEL_TRI * hp.local_verts      # rank 0's launch
        achieved = flops_per_launch / (hp.vv_kernel_ms * 1e-3) / 1e12
        fill_bytes = 0.25 * G + 8.0 * hp.nfluid + 48.0 * nbe + 16.0 * nvert + 4.0 * hp.ncell3
        key = "config%d" % args.config
        nv_ncu = NCU.get(key, {}).get("k_vert_volume", {}) if world == 1 and args.size is None else {}
        nf_ncu = NCU.get(key, {}).get("fill_resolve", {}) if world == 1 and args.size is None else {}
        sorted_gb = (16.0 * (nvert + nbe) + 36.0 * nbe) / 1e9
        line = {
            "metric": "end-to-end preprocess triangles/s (hot path)", "value": value, "unit": "triangles/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "baseline_config_index": args.config - 1, "nbe": nbe, "nvert": nvert, "lattice": [int(x) for x in hp.dimg],
                       "nfluid": hp.nfluid, "l2": "inputs+outputs of every stage > 126 MB L2 (sorted geometry %.2f GB, bitfield %.3f GB)"
                       % (sorted_gb, G / 8e9), "parallelism": hp.parallelism, "setup_s": t_setup},
            "stages": {"preprocess_s": ms / args.steps * 1e-3,
                       "vert_volume_verts_per_s": nalive / (vv_ms * 1e-3), "fill_cells_per_s": G / (fill_ms * 1e-3),
                       "ms": {k: float(v) for k, v in zip(hp.STAGES, stage_ms)}, "fill_rounds": hp.fill_rounds},
            "check": check,
            "roofline": {"kernel": "calc_vert_volume = k_vv_fans + k_vv_rows<AX> + k_vv_lines<AX> per chunk of vertices (one C-ABI call)",
                         "bound": "fp32_alu", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp32_peak, "traffic": nv_ncu.get("dram_bytes"),
                         "issue_slot_frac": nv_ncu.get("issue_slot_frac"), "instr_per_vertex": nv_ncu.get("warp_instr_per_vertex"),
                         "executed_fma_frac": nv_ncu.get("fma_pipe_frac"),
                         "frac_note": "ALGORITHMIC flops of the reference's brute-force 41^3 quadrature over the measured FFMA peak; the kernels "
                                      "remove > 95 % of them by exact culling (monotone threshold search), hence frac > 1.  The machine view is "
                                      "issue_slot_frac / instr_per_vertex (warp instructions) / executed_fma_frac from the whole-pass ncu capture of "
                                      "this workload (profiles/ncu_summary.json): the honest bound is instruction issue",
                         "peak_source": "FFMA micro-kernel measured in this run (cx_measure_fp32_peak); nominal 148*128*2*1.965 GHz = 74.4",
                         "flops_per_vertex": VOXELS * T * FLOP_PER_VOXEL_TRI, "mean_trisize": T,
                         "hbm_frac": (hp.vv_bytes / (hp.vv_kernel_ms * 1e-3) / 1e9) / hbm_peak if hp.vv_bytes else None},
            "roofline_fill": {"kernel": "whole geometry fill (k_seg_prep, k_resolve_cross + k_resolve_inplane, k_fill_tiles rounds)", "bound": "hbm",
                              "achieved": fill_bytes / (fill_ms * 1e-3) / 1e9,
                              "peak": hbm_peak, "unit": "GB/s", "frac": fill_bytes / (fill_ms * 1e-3) / 1e9 / hbm_peak,
                              "traffic": nf_ncu.get("dram_bytes"), "peak_source": hbm_src, "bytes": fill_bytes,
                              "issue_slot_frac": nf_ncu.get("issue_slot_frac"), "resolve_ms_under_ncu": nf_ncu.get("duration_ms_under_ncu"),
                              "frac_note": "SURVEY.md §8d bytes (0.25 B/cell + 8 B/filled cell + one pass over the geometry); traffic = DRAM bytes of "
                                           "the directed tests and their preparation (ncu).  The fill is bound by the latency of the propagation "
                                           "rounds and the directed tests near walls (instruction issue), not by HBM"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if not args.no_cpu and world == 1:
            B, kind, cores = cpu_backend()
            sample = wl.cpu_sample(args.config, args.steps + args.warmup)
            dt, nfl, nv_s = cpu_pass(B, sample)
            line["cpu_baseline"] = {"value": sample[2].shape[0] / dt, "unit": "triangles/s", "cores": cores, "kind": kind,
                                    "sample": sample_text(sample, nv_s) + " (%.1f s)" % dt}
        if not args.no_reference_cuda and not args.no_cpu and world == 1:
            line["reference_cuda"] = reference_cuda_leg()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        ctx.dist_finalize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
