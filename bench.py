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
        p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
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
SEED = 20261018          # counter-based synthetic vector (cdmft_lanc_ed_b200/synth.py): v(i) depends on the GLOBAL index only
NVLINK_GBS = 900.0       # NVLink 5 per direction and GPU (B200_PROFILING.md)
KINDS = {0: "column_pass", 1: "row_pass", 2: "transpose_pack_unpack", 3: "exchange_exposed", 6: "exchange_exposed_back", 5: "barrier"}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from cdmft_lanc_ed_b200 import models, synth
    from cdmft_lanc_ed_b200 import ed_hamiltonian as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback; use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        E.ed_set_MpiComm(local)
    else:
        E.ed_init(local)
    E.set_stream(torch.cuda.current_stream().cuda_stream)
    for kv in args.opt:
        E.set_option(kv.split("=")[0], int(kv.split("=")[1]))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(xs):
        t = torch.tensor(list(xs), dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    peaks = _load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"))
    peak = peaks["hbm_gbs"] if peaks and "hbm_gbs" in peaks else 6650.0
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks and "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"

    class Sector:
        """one workload sector on the device(s): build, counter-based vector, timed H x v, per-kind timers, checksum"""

        def __init__(self, workload, sparse):
            builder, margs, (nup, ndw) = WORKLOADS[workload]
            self.workload, self.sparse = workload, sparse
            self.mdl = getattr(models, builder)(*margs)
            E.ed_set_model(self.mdl)
            self.isec = models.get_sector(self.mdl.ns, nup, ndw)
            self.dim, self.dimup, self.dimdw = E.getDim(self.isec)
            t0 = time.perf_counter()
            self.nloc = E.build_Hv_sector(self.isec, sparse)
            if world > 1 and not args.no_ipc:
                E.ipc_exchange()  # CUDA-IPC windows: copy-engine exchange over NVLink
            self.build_s = time.perf_counter() - t0
            from cdmft_lanc_ed_b200 import shard_plan as sp
            self.off = sum(sp.vecdim(self.dimup, self.dimdw, world, r) for r in range(rank)) if world > 1 else 0
            self.scale = synth.default_scale(self.dim)
            self.v = synth.counter_vec_torch(self.off, self.nloc, SEED, self.scale)
            self.hv = torch.empty_like(self.v)
            small = 16 * self.dim <= 4 * 126e6
            self.flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda") if small else None

        def step(self):
            if self.flush is not None:
                self.flush.zero_()
            E.spHtimesV_p(self.nloc, self.v, self.hv)

        def time_hxv(self, steps, warmup):
            for _ in range(max(warmup, 3)):
                self.step()
            barrier()
            l0 = E.launch_count()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            if self.flush is not None:  # time the H x v only, not the flush
                ms = 0.0
                for _ in range(steps):
                    self.flush.zero_()
                    ev[0].record()
                    E.spHtimesV_p(self.nloc, self.v, self.hv)
                    ev[1].record()
                    torch.cuda.synchronize()
                    ms += ev[0].elapsed_time(ev[1])
            else:
                ev[0].record()
                for _ in range(steps):
                    E.spHtimesV_p(self.nloc, self.v, self.hv)
                ev[1].record()
                barrier()
                ms = ev[0].elapsed_time(ev[1])
            barrier()
            launches = E.launch_count() - l0
            return allmax(ms) / steps, launches

        def kinds(self, steps):
            """per-kind CUDA-event durations (events inside the library, on the launching stream)"""
            E.set_option("profile", 1)
            for k in KINDS:
                E.profile_query(k)
            n = min(steps, 10)
            for _ in range(n):
                self.step()
            out = {}
            for k, nm in KINDS.items():
                tot, cnt = E.profile_query(k)
                if cnt:
                    out[nm] = {"ms_per_step": tot / n, "launches_per_step": cnt / n}
            E.set_option("profile", 0)
            return out

        def checksum(self):
            """(Re<v,Hv>, |Hv|^2) over all ranks: the same numbers at every GPU count (counter-based v)"""
            torch.cuda.synchronize()
            E.spHtimesV_p(self.nloc, self.v, self.hv)
            torch.cuda.synchronize()
            a = float(torch.vdot(self.v, self.hv).real) if self.nloc else 0.0
            b = float(torch.vdot(self.hv, self.hv).real) if self.nloc else 0.0
            return allsum([a, b])

        def close(self):
            if E.spHtimesV_p is not None:
                E.delete_Hv_sector()
            self.v = self.hv = self.flush = None
            torch.cuda.empty_cache()

    def nvlink(sec, ms_step, kern):
        if world == 1:
            return None
        elem = 16
        out_bytes = 2.0 * (world - 1) / world * sec.nloc * elem  # two transposes, everything but the own block leaves
        return {"bytes_out_per_gpu_per_hxv": out_bytes, "peak_GBps_per_direction": NVLINK_GBS,
                "floor_ms": out_bytes / (NVLINK_GBS * 1e9) * 1e3,
                "frac_of_step": out_bytes / (ms_step * 1e-3) / 1e9 / NVLINK_GBS,
                "exposed_exchange_ms": ((kern.get("exchange_exposed") or {}).get("ms_per_step") or 0.0) + ((kern.get("exchange_exposed_back") or {}).get("ms_per_step") or 0.0),
                "barrier_ms": (kern.get("barrier") or {}).get("ms_per_step"),
                "note": "copy-engine exchange: the DMA copies overlap the column passes; exposed = what the compute stream still waits for its own outgoing copies; barrier = the two stream-ordered barriers per product, waiting for the slowest rank included"}

    def gs_lanczos(sec, niter, tol, fixed=None):
        """sp_lanc_eigh semantics from the constant start vector; warm call reported (the first call allocates)"""
        out = {}
        for rep in range(2):
            vec = torch.zeros(sec.nloc, dtype=torch.complex128, device="cuda")
            barrier()
            t0 = time.perf_counter()
            # fixed: exactly `fixed` iterations (ncheck past nitermax: the energy test never fires)
            e0, nit, _, _ = E.sp_lanc_eigh(vec, niter if fixed is None else fixed, tol if fixed is None else 1e-300,
                                           10 if fixed is None else fixed + 1)
            barrier()
            dt = allmax(time.perf_counter() - t0)
            out["first_call_seconds" if rep == 0 else "seconds"] = dt
            del vec
        real = sec.mdl.is_real and (world == 1 or sec.dimup % 2 == 0)
        out.update({"iterations": nit, "e0": e0, "threshold": tol if fixed is None else 0.0, "nitermax": niter if fixed is None else fixed,
                    "start": "constant 1/sqrt(Dim)",
                    "vectors": ("real (8 B) -- H and start vector real" + ("" if world == 1 else ", sharded: paired-row view")) if real else "complex(8)",
                    "krylov_vectors_kept_in_hbm": "as many as fit (option lanczos_store); the rest recomputed from the last two"})
        return out

    clk_path = os.path.join(ROOT, "gpurun_out", f"clocks_rank{rank}.csv")
    os.makedirs(os.path.dirname(clk_path), exist_ok=True)
    sampler = _clock_sampler_start(clk_path) if rank == 0 else (None, None)

    # ================= headline: args.workload (K3), SPARSE unless --direct =================
    sparse = not args.direct
    sec = Sector(args.workload, sparse)
    dim, nloc = sec.dim, sec.nloc
    ms_step, launches = sec.time_hxv(args.steps, args.warmup)
    value = 32.0 * dim / (ms_step * 1e-3) / 1e9
    kern = sec.kinds(args.steps)
    clocks = _clock_sampler_stop(*sampler, clk_path, gpu_index=local) if rank == 0 else None
    chk = sec.checksum()

    # ---- e2e: host buffers through the C ABI (H2D + kernels + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        vh = torch.empty(nloc, dtype=torch.complex128, pin_memory=True)
        hh = torch.empty(nloc, dtype=torch.complex128, pin_memory=True)
        vh.copy_(sec.v)
        E.spHtimesV_p(nloc, vh, hh)  # warm (allocates the staging buffers)
        nst = min(args.steps, 10)
        barrier()
        t0 = time.perf_counter()
        for _ in range(nst):
            E.spHtimesV_p(nloc, vh, hh)  # synchronous on return for host pointers
        barrier()
        dt = allmax(time.perf_counter() - t0)
        ok = bool(torch.equal(hh.cuda(), sec.hv))
        e2e = {"value": 32.0 * dim * nst / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": 16 * nloc * world,
               "d2h_bytes_per_step": 16 * nloc * world, "ms_per_step": dt / nst * 1e3, "steps": nst,
               "matches_device_path": ok, "call": "cdmft_b200_hxv64(nloc, host v, host hv) with pinned host buffers"}
        del vh, hh

    # ---- ground-state Lanczos seconds
    gs = None
    if not args.no_lanczos:
        gs = gs_lanczos(sec, args.lanc_niter, args.lanc_tol)
        gs["hxv_calls"] = "iterations + those recomputed for the eigenvector (0 when every Krylov vector fits in HBM)"
        fx = gs_lanczos(sec, 0, 0.0, fixed=100)
        gs["fixed_100_iterations"] = {"seconds": fx["seconds"], "iterations": fx["iterations"], "e0": fx["e0"]}
        fx_path = os.path.join(ROOT, "tests", "golden", "k3_gs_oracle.json")  # the CPU oracle's run of the same ground-state problem
        ofx = _load_json(fx_path) if args.workload == "K3" and args.lanc_tol == 1e-12 else None
        if ofx:
            gs["e0_oracle"] = ofx["e0"]
            gs["abs_e0_minus_oracle"] = abs(gs["e0"] - ofx["e0"])
            gs["iterations_oracle"] = ofx["iterations"]
            gs["oracle_seconds"] = ofx["oracle_seconds"]
            gs["oracle_note"] = f"sp_lanc_eigh restatement on {ofx['oracle_cores']} host cores (tests/golden/make_k3_gs.py; committed fixture, not run here)"
        if world == 1:  # the reference's default LANC_METHOD (sp_eigh, ED_DIAG.f90:150-170): device-resident thick-restart Lanczos
            try:
                t0 = time.perf_counter()
                we, _, inf = E.sp_eigh_device(2, nblock=20, nitermax=512, tol=args.lanc_tol,
                                              basis=torch.zeros((2, nloc), dtype=torch.complex128, device="cuda"))
                gs["sp_eigh_device"] = {"seconds": time.perf_counter() - t0, "neigen": 2, "nblock": 20, "tol": args.lanc_tol,
                                        "eig_values": [float(x) for x in we], "nconv": inf["nconv"], "hxv_calls": inf["nmatvec"],
                                        "note": "cdmft_b200_eigh: Krylov basis, full reorthogonalisation and restarts on the device"}
            except Exception as ex:  # a sub-record must not take the headline down
                gs["sp_eigh_device"] = {"error": str(ex)[:200]}
        if not args.no_e2e and world == 1:  # host start vector -> lanczos_gs -> host eigenvector
            hv0 = torch.zeros(nloc, dtype=torch.complex128, pin_memory=True)
            t0 = time.perf_counter()
            e0h, nith, _, _ = E.sp_lanc_eigh(hv0, args.lanc_niter, args.lanc_tol)
            gs["e2e_host_buffers_seconds"] = time.perf_counter() - t0
            gs["e2e_host_buffers_e0"] = e0h
            del hv0

    # ---- roofline: B_alg = 32 B x Dim per H x v against the measured HBM copy bandwidth
    prof = _load_json(os.path.join(ROOT, "profiles", "r2_ncu_summary.json")) or _load_json(os.path.join(ROOT, "profiles", "r1_ncu_summary.json")) or {}
    model_bytes = {"column_pass": 32.0, "row_pass": 48.0}
    per_kernel = {}
    tot_kern = sum(x["ms_per_step"] for x in kern.values()) or 1.0
    for k, x in kern.items():
        if k in model_bytes and x["ms_per_step"] > 0:
            # sharded runs launch the column pass twice (Hup on v, Hdw on vt): bytes per launch = model x shard states
            ach = model_bytes[k] * nloc * max(1.0, round(x["launches_per_step"])) / (x["ms_per_step"] * 1e-3) / 1e9 if world == 1 else \
                model_bytes[k] * nloc * 2.0 / (x["ms_per_step"] * 1e-3) / 1e9
            per_kernel[k] = {"algorithmic_bytes_per_state": model_bytes[k], "ms_per_step": x["ms_per_step"], "achieved": ach, "frac": ach / peak,
                             "share_of_step": x["ms_per_step"] / tot_kern,
                             "traffic_ncu_static": (prof.get(args.workload, {}).get(k, {}) or {}).get("dram_bytes_per_launch") if world == 1 else None}
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_step"]) if per_kernel else None
    names = {"column_pass": "k_colres (column-resident shared-memory kernel; k_colblk / k_colpass when a column does not fit)",
             "row_pass": "k_rowpass_rb (L2-slab row pass)"}
    roofline = {"bound": "hbm", "achieved": value / world, "peak": peak, "unit": "GB/s", "frac": value / world / peak,
                "definition": "SURVEY 8(d): 32 B x Dim / t_step / (n_gpus x peak) -- read v once, write Hv once",
                "peak_source": peak_src, "dominant_kernel": names.get(dom, dom),
                "traffic": per_kernel.get(dom, {}).get("traffic_ncu_static") if dom else None,
                "traffic_source": "dram__bytes_read+write of that kernel from the ncu --set full capture under profiles/ (static, not measured in this run)",
                "kernels": per_kernel, "nvlink": nvlink(sec, ms_step, kern)}

    # ================= sub-records: the other BASELINE configs (few steps each) =================
    sub = {}
    if not args.no_sub:
        ns_ = min(args.steps, 5)
        if sparse:
            sec.close()
            d = Sector(args.workload, False)  # BASELINE config 3: DIRECT (matrix-free) H x v
            ms_d, _ = d.time_hxv(ns_, 3)
            cd = d.checksum()
            sub[f"{args.workload}_direct"] = {"ms_per_step": ms_d, "GBps": 32.0 * d.dim / ms_d / 1e6, "frac": 32.0 * d.dim / ms_d / 1e6 / world / peak,
                                              "checksum": cd, "kernels": d.kinds(ns_)}
            d.close()
        else:
            sec.close()
        if args.workload == "K3" and world == 1:  # BASELINE config 4: complex Hamiltonian (BHZ), Ns=16
            k4 = Sector("K4", True)
            ms4, _ = k4.time_hxv(ns_, 3)
            sub["K4"] = {"ms_per_step": ms4, "GBps": 32.0 * k4.dim / ms4 / 1e6, "frac": 32.0 * k4.dim / ms4 / 1e6 / peak,
                         "checksum": k4.checksum(), "kernels": k4.kinds(ns_)}
            if not args.no_lanczos:
                g4 = gs_lanczos(k4, args.lanc_niter, args.lanc_tol)
                sub["K4"]["gs_lanczos"] = {k: g4[k] for k in ("seconds", "iterations", "e0", "vectors")}
            k4.close()
        if world >= 8 and not args.no_k5:  # BASELINE config 5 / north-star target: Ns=18 half filling on the whole box
            sub["K5"] = _k5_record(E, Sector, gs_lanczos, nvlink, args, world, rank, peak, allmax, barrier)
    else:
        sec.close()

    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu:
        r = _cpu_port_hxv(args.workload, budget_s=40.0, steps=1, warmup=0)
        cpu = {"value": r["value"], "unit": "GB/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        if gs:
            gs["cpu_port_seconds_estimate"] = gs["iterations"] * 2 * r["ms_per_step"] * 1e-3 * (dim / r["dim"])
            gs["cpu_port_note"] = "two-pass sp_lanc_eigh on the CPU port = 2 x iterations H x v at the measured CPU rate (vector ops not counted)"
    E.ed_finalize()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        line = {
            "metric": "hxv_algorithmic_GBps", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "c128", "data": f"synthetic (counter-based vector, seed {SEED})",
            "config": _config(args.workload, world, sparse),
            "hxv_per_s": 1e3 / ms_step, "gpu_launches": launches, "checksum": {"re_v_hv": chk[0], "hv_norm2": chk[1]},
            "kernels": kern, "e2e": e2e, "roofline": roofline,
            "cpu_baseline": cpu, "gs_lanczos": gs, "configs": sub, "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    return 0


def _k5_record(E, Sector, gs_lanczos, nvlink, args, world, rank, peak, allmax, barrier):
    """Ns=18 half-filled sector (Dim 2 363 904 400; the reference's int32 getDim overflows, ED_SETUP.f90:321):
    H x v, NVLink share, ground-state Lanczos and a sampled-row check against the oracle."""
    import numpy as np
    import torch
    import torch.distributed as dist
    k5 = Sector("K5", True)
    ns_ = min(args.steps, 5)
    ms5, _ = k5.time_hxv(ns_, 3)
    kern5 = k5.kinds(ns_)
    rec = {"dim": k5.dim, "nloc_rank0": k5.nloc, "build_s": k5.build_s, "ms_per_step": ms5, "GBps": 32.0 * k5.dim / ms5 / 1e6,
           "frac_of_aggregate_hbm": 32.0 * k5.dim / ms5 / 1e6 / world / peak, "kernels": kern5, "nvlink": nvlink(k5, ms5, kern5),
           "checksum": k5.checksum()}
    # sampled rows: every rank reads a few hundred rows of ITS shard of Hv; rank 0 evaluates the same rows with the
    # oracle from the index function alone (no 38 GB host vector)
    rng = np.random.default_rng(77 + rank)
    loc = np.unique(rng.integers(0, max(k5.nloc, 1), size=300)) if k5.nloc else np.zeros(0, dtype=np.int64)
    vals = k5.hv[torch.from_numpy(loc).cuda()].cpu().numpy() if k5.nloc else np.zeros(0, dtype=np.complex128)
    parts = [None] * world
    dist.all_gather_object(parts, (loc + k5.off, vals))
    if rank == 0:
        from oracle import edo  # checker only
        edo.lib().edo_set_num_threads(os.cpu_count() or 1)
        orc = edo.Oracle(k5.mdl)
        orc.build_hv_sector(k5.isec, edo.DIRECT_SERIAL)
        rows = np.concatenate([p[0] for p in parts])
        got = np.concatenate([p[1] for p in parts])
        ref = orc.hxv_rows_counter(rows, SEED, k5.scale)
        orc.delete_hv_sector()
        rec["sampled_rows"] = {"rows": int(rows.size), "max_abs_err": float(np.abs(got - ref).max()),
                               "max_rel_err": float(np.abs(got - ref).max() / np.abs(ref).max()),
                               "oracle": "edo_hxv_rows_counter (pull form of the direct products on the counter-based vector)"}
    if not args.no_lanczos:
        g5 = gs_lanczos(k5, args.lanc_niter, args.lanc_tol)
        rec["gs_lanczos"] = {k: g5[k] for k in ("first_call_seconds", "seconds", "iterations", "e0", "vectors", "threshold")}
    k5.close()
    return rec


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
    ap.add_argument("--no-sub", action="store_true", help="skip the sub-records (DIRECT mode, K4, and K5 at 8 GPUs)")
    ap.add_argument("--no-k5", action="store_true")
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
