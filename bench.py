#!/usr/bin/env python
"""bench.py -- H x v throughput of the CDMFT-LANC-ED hot path on B200 (driver contract).

A "step" is one Hamiltonian-times-vector product over the whole sector vector
(spHtimesV_p, ED_VARS_GLOBAL.f90:72-78) of the workload BASELINE.json's metric is quoted on:
K3 = cdn_hm_2dsquare 2x2 plaquette, Nbath=3 (Ns=16), sector (8,8), Dim 165 636 900, complex(8).

  metric   hxv_algorithmic_GBps = B_alg * steps / time,  B_alg = 2*16*Dim bytes
  value    inputs resident in HBM, library kernels only (CUDA events, max over ranks)
  e2e      same metric through the C-ABI call with HOST buffers (pinned): H2D + kernels + D2H per step
  N > 1    the vector is sharded along Ndw exactly like direct_mpi (ED_HAMILTONIAN.f90:92-105), the
           two transposes per H x v are NCCL all-to-alls; total work is fixed -> "scaling": "strong"

`--impl reference` times the reference algorithm's CPU restatement (oracle/, spMatVec_mpi_main with
the MPI ranks simulated as OpenMP threads on all host cores; the Fortran reference itself cannot be
built in this image: no gfortran / MPI / SciFortran).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model builder, args, (nup, ndw))
    "K1": ("hm2x2", (1,), (4, 4)),
    "K2": ("hm2x2", (2,), (6, 6)),
    "K3": ("hm2x2", (3,), (8, 8)),
    "K4": ("bhz2", (3,), (8, 8)),
    "K5": ("hm_ns18", (), (9, 9)),  # Ns=18, Dim 2 363 904 400: needs >= 2 GPUs (int64 global index; the reference overflows)
}


def _clock_sampler_start(path):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        f = open(path, "w")
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                             stdout=f, stderr=subprocess.DEVNULL)
        return p, f
    except Exception:
        return None, None


def _clock_sampler_stop(p, f, path, gpu_index=0):
    if p is None:
        return None
    p.terminate()
    try:
        p.wait(timeout=5)
    except Exception:
        p.kill()
    f.close()
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    try:
        for line in open(path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9 or not parts[0].isdigit() or int(parts[0]) != gpu_index:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
    except Exception:
        return None
    if not sm:
        return None
    busy = [x for x in sm if x > 0.5 * max(sm)] or sm
    return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def _load_json(path):
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------
def _cpu_port_hxv(workload, budget_s, steps, warmup, full_sector_first=True):
    """Times `steps` H x v of the CPU restatement (after `warmup`) on a sector of the workload's model
    sized so that the whole run fits `budget_s`.  Returns dict(value GB/s, ms_per_step, cores, sample)."""
    import numpy as np
    from cdmft_lanc_ed_b200 import models
    from oracle import edo
    builder, args, (nup, ndw) = WORKLOADS[workload]
    mdl = getattr(models, builder)(*args)
    ns = mdl.ns
    cores = os.cpu_count() or 1
    edo.lib().edo_set_num_threads(cores)
    orc = edo.Oracle(mdl)
    # calibrate ns/state on a small sector of the same model
    cal = models.get_sector(ns, 2, ndw)
    orc.build_hv_sector(cal, edo.SPARSE_MPI, cores)
    v = np.ones(orc.dim, dtype=np.complex128)
    orc.hxv(v)
    t0 = time.perf_counter()
    orc.hxv(v)
    rate = (time.perf_counter() - t0) / orc.dim  # s per state
    orc.delete_hv_sector()
    # largest (n, ndw) sector, n <= nup, whose (steps+warmup) products (+ ~equal build time) fit the budget
    from math import comb
    pick = 2
    for n in range(nup, 1, -1):
        dim = comb(ns, n) * comb(ns, ndw)
        if dim * rate * (steps + warmup + 1.0) * 1.6 <= budget_s:
            pick = n
            break
    isec = models.get_sector(ns, pick, ndw)
    tb = time.perf_counter()
    orc.build_hv_sector(isec, edo.SPARSE_MPI, cores)
    tb = time.perf_counter() - tb
    dim = orc.dim
    rng = np.random.default_rng(12345)
    v = (rng.standard_normal(dim) + 1j * rng.standard_normal(dim)).astype(np.complex128)
    for _ in range(warmup):
        orc.hxv(v)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.hxv(v)
    dt = time.perf_counter() - t0
    orc.delete_hv_sector()
    orc.close()
    return {
        "value": 32.0 * dim * steps / dt / 1e9, "ms_per_step": dt / steps * 1e3, "cores": cores, "dim": dim,
        "sample": f"{steps} H x v of model {mdl.name} sector ({pick},{ndw}) dim={dim} "
                  f"({'the full workload sector' if pick == nup else 'bounded sample of the workload sector (%d,%d)' % (nup, ndw)}), "
                  f"spMatVec_mpi_main restatement (oracle/ed_oracle.c), {cores} simulated MPI ranks = OpenMP threads; "
                  f"sector build {tb:.1f} s not timed",
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = _cpu_port_hxv(args.workload, budget_s=150.0, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "hxv_algorithmic_GBps", "value": r["value"], "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "c128", "data": "synthetic",
        "config": _config(args.workload, args.gpus, sparse=True),
        "cpu_baseline": {"value": r["value"], "unit": "GB/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the Fortran reference cannot be compiled here (no gfortran/MPI/SciFortran): kind=port is the C "
                "restatement of its MPI sparse mat-vec; throughput is size-normalised (32 B per state)",
    }
    print(json.dumps(line), flush=True)
    return 0


def _config(workload, ngpus, sparse):
    builder, margs, sec = WORKLOADS[workload]
    from math import comb
    from cdmft_lanc_ed_b200 import models
    mdl = getattr(models, builder)(*margs)
    dim = comb(mdl.ns, sec[0]) * comb(mdl.ns, sec[1])
    return {
        "workload": f"{workload}: {mdl.name} Ns={mdl.ns} sector ({sec[0]},{sec[1]}) Dim={dim} complex(8), "
                    f"one H x v per step, ed_sparse_H={'T' if sparse else 'F'}",
        "dim": dim, "bytes_alg_per_step": 32 * dim,
        "l2": f"vector {16 * dim / 1e6:.0f} MB per pass vs 126 MB L2 (inputs larger than L2, no flush needed)"
              if 16 * dim > 4 * 126e6 else "L2 flushed between steps by a 512 MB memset",
        "sharding": "none (single rank)" if ngpus == 1 else f"Ndw split over {ngpus} ranks (ED_HAMILTONIAN.f90:92-105), distributed transposes over NVLink",
    }


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from cdmft_lanc_ed_b200 import models
    from cdmft_lanc_ed_b200 import ed_hamiltonian as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback; use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        E.ed_set_MpiComm(local)
    else:
        E.ed_init(local)
    E.set_stream(torch.cuda.current_stream().cuda_stream)
    for kv in args.opt:
        E.set_option(kv.split("=")[0], int(kv.split("=")[1]))

    builder, margs, (nup, ndw) = WORKLOADS[args.workload]
    mdl = getattr(models, builder)(*margs)
    E.ed_set_model(mdl)
    isec = models.get_sector(mdl.ns, nup, ndw)
    dim = E.getDim(isec)[0]
    sparse = not args.direct
    nloc = E.build_Hv_sector(isec, sparse)
    if world > 1 and not args.no_ipc:
        E.ipc_exchange()  # peer-memory transposes (CUDA IPC windows over NVLink)
    small = 16 * dim <= 4 * 126e6
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda") if small else None

    # seeded synthetic vector (one stream per rank)
    g = torch.Generator(device="cuda")
    g.manual_seed(12345 + rank)
    v = torch.randn(nloc, 2, dtype=torch.float64, device="cuda", generator=g).view(-1)
    v = torch.view_as_complex(v.view(nloc, 2)).contiguous()
    hv = torch.empty_like(v)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        if flush is not None:
            flush.zero_()
        E.spHtimesV_p(nloc, v, hv)

    clk_path = os.path.join(ROOT, "gpurun_out", f"clocks_rank{rank}.csv")
    os.makedirs(os.path.dirname(clk_path), exist_ok=True)
    sampler = _clock_sampler_start(clk_path) if rank == 0 else (None, None)
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    l0 = E.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    if small:  # time the H x v only, not the flush
        tot = 0.0
        for _ in range(args.steps):
            flush.zero_()
            ev[0].record()
            E.spHtimesV_p(nloc, v, hv)
            ev[1].record()
            torch.cuda.synchronize()
            tot += ev[0].elapsed_time(ev[1])
        ms = tot
    else:
        ev[0].record()
        for _ in range(args.steps):
            E.spHtimesV_p(nloc, v, hv)
        ev[1].record()
        barrier()
        ms = ev[0].elapsed_time(ev[1])
    barrier()
    launches = E.launch_count() - l0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_step = ms / args.steps
    value = 32.0 * dim / (ms_step * 1e-3) / 1e9

    # ---- per-kernel durations (separate profiled loop, CUDA events inside the library, launch stream)
    E.set_option("profile", 1)
    for _ in range(min(args.steps, 10)):
        step()
    kinds = {0: "column_pass", 1: "row_pass", 2: "transpose_pack_unpack", 3: "nccl_all_to_all"}
    kern = {}
    nprof = min(args.steps, 10)
    for k, nm in kinds.items():
        tot, n = E.profile_query(k)
        if n:
            kern[nm] = {"ms_per_step": tot / nprof, "launches_per_step": n / nprof}
    E.set_option("profile", 0)
    # clocks were sampled (nvidia-smi -lms 100) from the first warm-up step to the end of the profiled loop
    clocks = _clock_sampler_stop(*sampler, clk_path, gpu_index=local) if rank == 0 else None

    # ---- e2e: host buffers through the C ABI (H2D + kernels + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        vh = torch.empty(nloc, dtype=torch.complex128, pin_memory=True)
        hh = torch.empty(nloc, dtype=torch.complex128, pin_memory=True)
        vh.copy_(v)
        E.spHtimesV_p(nloc, vh, hh)  # warm (allocates the staging buffers)
        nst = min(args.steps, 10)
        barrier()
        t0 = time.perf_counter()
        for _ in range(nst):
            E.spHtimesV_p(nloc, vh, hh)  # synchronous on return for host pointers
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        ok = bool(torch.allclose(hh.cuda(), hv, rtol=0, atol=0))
        e2e = {"value": 32.0 * dim * nst / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": 16 * nloc * world,
               "d2h_bytes_per_step": 16 * nloc * world, "ms_per_step": dt / nst * 1e3, "steps": nst,
               "matches_device_path": ok, "call": "cdmft_b200_hxv64(nloc, host v, host hv) with pinned host buffers"}
        del vh, hh

    # ---- ground-state Lanczos seconds (sp_lanc_eigh semantics, constant start vector)
    gs = None
    if not args.no_lanczos:
        vec = torch.zeros(nloc, dtype=torch.complex128, device="cuda")
        barrier()
        t0 = time.perf_counter()
        e0, nit, _, _ = E.sp_lanc_eigh(vec, args.lanc_niter, args.lanc_tol)
        barrier()
        gs = {"seconds": time.perf_counter() - t0, "iterations": nit, "hxv_calls": 2 * nit, "e0": e0,
              "threshold": args.lanc_tol, "nitermax": args.lanc_niter, "start": "constant 1/sqrt(Dim)",
              "vectors": ("real (8 B) -- H and start vector real" + ("" if world == 1 else ", sharded: paired-row view"))
                         if (mdl.is_real and (world == 1 or E.getDim(isec)[1] % 2 == 0)) else "complex(8)"}
        del vec

    peaks = _load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"))
    peak = peaks["hbm_gbs"] if peaks and "hbm_gbs" in peaks else 6650.0
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks and "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    prof = _load_json(os.path.join(ROOT, "profiles", "r1_ncu_summary.json")) or {}
    roofline = None
    # dominant kernel of the step = the kind with the largest share of the profiled loop.  Algorithmic bytes per
    # launch: column pass reads its operand once and writes its output once (32 B/state of
    # the shard, 16 B/state in real mode is not timed here); row pass reads v and read-modify-writes Hv (48 B/state).
    cand = {k: kern[k] for k in ("column_pass", "row_pass") if k in kern}
    if cand:
        dom = max(cand, key=lambda k: cand[k]["ms_per_step"])
        nl = cand[dom]["launches_per_step"]
        dur = cand[dom]["ms_per_step"] / max(nl, 1)
        per_state = 32.0 if dom == "column_pass" else 48.0
        bytes_launch = per_state * nloc
        ach = bytes_launch / (dur * 1e-3) / 1e9
        names = {"column_pass": "column pass (k_colres: column-resident shared-memory kernel; k_colpass when a column does not fit)",
                 "row_pass": "row pass (k_rowpass_rb)"}
        roofline = {"bound": "hbm", "kernel": names[dom], "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "peak_source": peak_src,
                    "traffic": (prof.get(args.workload, {}).get(dom, {}) or {}).get("dram_bytes_per_launch") if world == 1 else None,
                    "algorithmic_bytes_per_launch": bytes_launch, "algorithmic_bytes_per_state": per_state, "avg_launch_ms": dur,
                    "share_of_step": cand[dom]["ms_per_step"] / sum(x["ms_per_step"] for x in kern.values()),
                    "other_kernels": {k: {"achieved": (32.0 if k == "column_pass" else 48.0) * nloc * x["launches_per_step"] / (x["ms_per_step"] * 1e-3) / 1e9,
                                          "frac": (32.0 if k == "column_pass" else 48.0) * nloc * x["launches_per_step"] / (x["ms_per_step"] * 1e-3) / 1e9 / peak}
                                      for k, x in cand.items() if k != dom},
                    "step_frac_of_peak": value / world / peak}

    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu:
        E.delete_Hv_sector()
        torch.cuda.empty_cache()
        r = _cpu_port_hxv(args.workload, budget_s=40.0, steps=1, warmup=0)
        cpu = {"value": r["value"], "unit": "GB/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
    if E.spHtimesV_p is not None:
        E.delete_Hv_sector()
    E.ed_finalize()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        line = {
            "metric": "hxv_algorithmic_GBps", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "c128", "data": "synthetic",
            "config": _config(args.workload, world, sparse),
            "hxv_per_s": 1e3 / ms_step, "gpu_launches": launches, "kernels": kern, "e2e": e2e, "roofline": roofline,
            "cpu_baseline": cpu, "gs_lanczos": gs, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="K3", choices=list(WORKLOADS))
    ap.add_argument("--direct", action="store_true", help="ed_sparse_H=F (matrix-free kernels)")
    ap.add_argument("--no-ipc", action="store_true", help="N>1: NCCL all-to-all transposes instead of peer-memory stores")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE", help="library option (cdmft_b200_set_option), repeatable")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-lanczos", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--lanc-niter", type=int, default=512)
    ap.add_argument("--lanc-tol", type=float, default=1e-12)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            # convenience: re-launch under torchrun, one rank per GPU
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
            return subprocess.call(cmd)
        raise SystemExit(f"--gpus {args.gpus} != WORLD_SIZE {world}")
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
