#!/usr/bin/env python
"""bench.py -- A+B Galerkin assembly throughput (structural nonzeros/s, A and B counted) on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
  * workload      BASELINE.json configs[2]: test_mesh_a.json, Orders(6,6), 6x global T (1,178,112 DoFs, 57,557,904 upper-
                  triangular entries per matrix), HierPoly / CurlCurl / L2Inner, GLQ [8,8], EXACT (bit-faithful) mode.
  * value         device-resident: plan + domain already in HBM, one step = sampler (K1) + integrator (K2) + scatter (K3)
                  into device CSR value arrays; CUDA events on the launch stream, max over ranks; the step is replayed from a CUDA graph.
  * roofline      dominant kernel of the headline step = K3 gather/scatter (HBM bound): algorithmic bytes = 16 B x nnz_upper + the
                  packed source map (SURVEY.md 8d budgets a 4 B index; the smaller figure the kernel really reads is used).
  * workloads     first-class second results where the integrator IS the step: `hp1m` (north_star target, >= 1M-DoF anisotropic
                  hp-mesh, every N) and `cfg3_dedupe0` (headline mesh without block dedupe, N = 1), each with an FP64-issue roofline
                  block (algorithmic lane-ops of the reference's per-pair quadrature / integrator time / measured DMUL+DADD peak);
                  `roofline_fp64` repeats hp1m's.
  * e2e           the reference-facing one-shot call with HOST buffers, every step: fem2d_galerkin_sample_gep_hcurl_multi on all N GPUs
                  from one process (host planner + symbolic + numeric + D2H + host expansion of rows/cols), wall clock.
  * cpu_baseline  the C++ oracle (restatement of the Rayon path; kind "port") on the SAME full workload, all host threads, once.
  * --impl reference   the same CPU restatement, full workload per step, as many steps as fit a 600 s budget (the Rust reference
                  cannot be built here: no rustc/cargo).
N > 1: strong scaling of the same meshes -- every rank owns a row block of the pattern (fem2d_plan_row_blocks_split), integrates what
its slots read and scatters its block; no data-path collective (halo tiles are recomputed).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # name: (recipe kwargs, glq)
    "cfg3": dict(mesh="a", order=6, levels=6, glq=8),        # BASELINE.json configs[2] (headline, ~1M DoFs)
    "cfg3_l5": dict(mesh="a", order=6, levels=5, glq=8),
    "cfg3_l4": dict(mesh="a", order=6, levels=4, glq=8),
    "cfg3_l3": dict(mesh="a", order=6, levels=3, glq=8),
    "cfg3_l2": dict(mesh="a", order=6, levels=2, glq=8),
    "cfg2": dict(mesh="b", order=8, levels=3, glq=12),       # BASELINE.json configs[1]
    "cfg4": dict(mesh="c4", glq=12),                         # BASELINE.json configs[3] (anisotropic, random p)
    # north_star target: >= 1M-DoF anisotropically hp-refined domain (cfg-4 recipe at 6 T-levels, 4 U/V rounds: 1,380,549 DoFs)
    "hp1m": dict(mesh="hp", glq=12),
}


WORKLOAD_TEXT = {
    "cfg3": "cfg3: test_mesh_a.json, Orders(6,6), 6x global T (BASELINE.json configs[2]), HierPoly/CurlCurl/L2Inner, GLQ [8,8]",
    "cfg2": "cfg2: test_mesh_b.json, Orders(8,8), 3x global T (BASELINE.json configs[1]), GLQ [12,12]",
    "cfg4": "cfg4: test_mesh_c.json, 4x T + 3 seeded U/V rounds, p in [2,10] (BASELINE.json configs[3]), GLQ [12,12]",
    "hp1m": "hp1m: test_mesh_c.json, 6x global T + 4 seeded anisotropic U/V rounds, random p in [2,10] per Elem (north_star target: >= 1M-DoF "
            "anisotropically hp-refined H(curl) domain; 1,380,549 DoFs), HierPoly/CurlCurl/L2Inner, GLQ [12,12]",
}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def _ncu_traffic(workload, dedupe, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of one k3_gather_kernel launch from the committed `ncu --set full` capture of this
    workload (profiles/ncu_traffic.json, written from the .ncu-rep by scripts/ncu_summary.py); None when no capture matches."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if world != 1 or not os.path.exists(p):
        return None
    with open(p) as f:
        t = json.load(f)
    e = t.get(f"k3_gather_kernel/{workload}/dedupe{int(dedupe)}")
    return None if e is None else float(e["dram_bytes_read"] + e["dram_bytes_write"])


def _ncu_traffic_k2(workload, dedupe, world):
    """DRAM bytes of the integrator's kernels (persistent grid + small-item grid) of one launch, from the committed ncu captures."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if world != 1 or not os.path.exists(p):
        return None
    with open(p) as f:
        t = json.load(f)
    tot, seen = 0.0, False
    for k in ("k2_ws_kernel", "k2_exact_kernel<4,64>"):
        e = t.get(f"{k}/{workload}/dedupe{int(dedupe)}")
        if e is not None:
            tot += e["dram_bytes_read"] + e["dram_bytes_write"]; seen = True
    return tot if seen else None


def build_product_domain(workload: str):
    import fem_2d_b200 as F
    import recipes
    w = WORKLOADS[workload]
    api = recipes.api("product")
    if w["mesh"] == "a":
        m = recipes.mesh_cfg3(api, levels=w["levels"], order=w["order"])
    elif w["mesh"] == "b":
        m = recipes.mesh_cfg2(api, levels=w["levels"], order=w["order"])
    elif w["mesh"] == "hp":
        m = recipes.mesh_hp1m(api)
    else:
        m = recipes.mesh_cfg4(api)
    return F.Domain.from_mesh(m)


def build_oracle_domain(workload: str):
    import oracle as O
    import recipes
    w = WORKLOADS[workload]
    api = recipes.api("oracle")
    if w["mesh"] == "a":
        m = recipes.mesh_cfg3(api, levels=w["levels"], order=w["order"])
    elif w["mesh"] == "b":
        m = recipes.mesh_cfg2(api, levels=w["levels"], order=w["order"])
    elif w["mesh"] == "hp":
        m = recipes.mesh_hp1m(api)
    else:
        m = recipes.mesh_cfg4(api)
    return O.Domain.from_mesh(m)


class ClockSampler:
    """Samples SM clocks / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False
        self._stop = threading.Event()
        self._t = None

    def sample(self):
        if not self.ok:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}
            for bit, nm in names.items():
                if r & bit:
                    self.reasons.add(nm)
        except Exception:
            pass

    def start(self, period_s=0.002):
        def loop():
            while not self._stop.is_set():
                self.sample()
                time.sleep(period_s)
        self._t = threading.Thread(target=loop, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        self.sample()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_port_run(workload: str, threads: int):
    """One assembly of `workload` by the C++ oracle (restatement of the Rayon path) with `threads` workers."""
    import oracle as O
    d = build_oracle_domain(workload)
    g = WORKLOADS[workload]["glq"]
    t0 = time.perf_counter()
    gep = O.galerkin_sample_gep_hcurl(d, [g, g], n_threads=threads)
    t1 = time.perf_counter()
    return dict(nnz=len(gep.rows), seconds=t1 - t0, integrate_s=gep.t_integrate, merge_s=gep.t_merge, dofs=d.num_dofs)


def run_reference(args):
    """CPU arm on the SAME config as the GPU arm: the full headline workload (BASELINE configs[2]) assembled by the C++ restatement of the
    Rayon path on all host threads.  One assembly takes about half a minute (the serial ordered-map merge of linalg.rs:74-79 dominates), so
    the number of repetitions is bounded by a wall-clock budget; `steps_timed` says how many were run."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    workload = args.workload
    budget = float(os.environ.get("FEM2D_REF_BUDGET_S", "600"))
    t_begin = time.perf_counter()
    results = []
    warm = None
    if args.warmup > 0:
        warm = cpu_port_run(workload, threads)          # also tells how many timed steps fit
    est = warm["seconds"] if warm else None
    for k in range(args.steps):
        if results and (time.perf_counter() - t_begin) + (est or results[-1]["seconds"]) > budget:
            break
        results.append(cpu_port_run(workload, threads))
        est = results[-1]["seconds"]
    sec = sum(r["seconds"] for r in results)
    nnz2 = sum(2 * r["nnz"] for r in results)
    value = nnz2 / sec
    line = {
        "impl": "reference", "metric": "assembly_nnz_per_s", "value": value, "unit": "nnz/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "steps_timed": len(results), "warmup_run": 1 if warm else 0, "ms_per_step": 1e3 * sec / len(results),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT.get(workload, workload), "mode": "exact", "n_dofs": results[0]["dofs"],
                   "nnz_upper_per_matrix": results[0]["nnz"], "nnz_counted": "2 x nnz_upper (A and B)",
                   "same_config_as_gpu_arm": workload == "cfg3",
                   "repetitions": f"{len(results)} timed full assemblies within a {budget:.0f} s budget (requested {args.steps})"},
        "cpu_baseline": {"value": value, "unit": "nnz/s", "cores": threads, "kind": "port",
                         "sample": f"full workload ({results[0]['nnz']} upper entries/matrix), C++ restatement of the Rayon path; integrate {results[0]['integrate_s']:.2f}s ({threads} threads) + serial merge {results[0]['merge_s']:.2f}s per step"},
        "e2e": {"value": value, "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "Rust reference cannot be built in this image (no rustc/cargo): CPU arm = C++ restatement (oracle), all host threads",
    }
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args):
    import numpy as np
    import torch
    import fem_2d_b200 as F

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or F.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device -- the fem_2d_b200 numeric path has no CPU fallback")
    workload = args.workload
    w = WORKLOADS[workload]
    mode = {"exact": F.MODE_EXACT, "sumfact": F.MODE_SUMFACT, "dmma": F.MODE_DMMA}[args.mode]
    domain = build_product_domain(workload)
    view = domain.view()
    glq = (F.gauss_quadrature_points(w["glq"]), F.gauss_quadrature_points(w["glq"]))
    dist = None
    nccl = None
    e2e = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        # The end-to-end leg comes first at N > 1: rank 0 drives all N GPUs from ONE process (fem2d_galerkin_sample_gep_hcurl_multi) while
        # the other ranks wait on a CPU-side (gloo) barrier and have not created their CUDA contexts yet -- a second process's context on a
        # GPU, even an idle one behind an NCCL barrier, makes the GPU switch contexts under rank 0's bursts (measured: 21 -> 36-54 ms).
        dist.init_process_group(backend="gloo")
        if not args.no_e2e:
            e2e = run_e2e(F, view, glq, mode, args, 0, rank, world, dist)
            if rank == 0:
                F.trim_cache()
            dist.barrier()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        nccl = dist.new_group(backend="nccl")

    def barrier():
        if dist is not None:
            dist.barrier(group=nccl)
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=nccl)
        return float(t.item())

    stream = torch.cuda.current_stream()
    peaks, peak_kind = _peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    fp64 = {}
    try:   # FP64 pipe denominators, measured on this GPU in this run (the EXACT integrator issues only non-fusable DMUL / DADD)
        fp64 = {"dfma_gflops": F.fp64_peak(local_rank, 0), "dmul_dadd_gflops": F.fp64_peak(local_rank, 1)}
    except Exception as ex:  # pragma: no cover
        fp64 = {"error": str(ex)}
    fp64_peak_gops = float(fp64.get("dmul_dadd_gflops", 18300.0))

    launch_kind = {}

    def timed(step, steps, warmup, sampler=None, tag="headline"):
        """`steps` steps between two CUDA events on the launch stream, barrier + synchronize on both sides, max over ranks -> ms per step.
        The step (sampler + integrator + scatter launches chained by programmatic dependent launch) is captured once into a CUDA graph and
        replayed: at N = 8 a step is 50 us of GPU work, and three launches through ctypes per step are then bounded by the host, not by
        the GPU.  Falls back to direct launches if the capture fails."""
        for _ in range(warmup):
            step(stream.cuda_stream)
        torch.cuda.synchronize()
        run = lambda: step(stream.cuda_stream)
        launch_kind[tag] = "direct"
        if not args.no_graph:
            try:
                cap = torch.cuda.Stream()
                cap.wait_stream(stream)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=cap):
                    step(torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                graph.replay(); torch.cuda.synchronize()
                run = graph.replay
                launch_kind[tag] = "cuda_graph"
            except Exception as ex:  # pragma: no cover
                launch_kind[tag] = f"direct (graph capture failed: {ex})"
                torch.cuda.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler is not None:
            sampler.start()
        e0.record(stream)
        for _ in range(steps):
            run()
        e1.record(stream)
        while not e1.query():
            time.sleep(0.0005)
        torch.cuda.synchronize()
        if sampler is not None:
            sampler.stop()
        ms = max_over_ranks(e0.elapsed_time(e1))
        barrier()
        return ms / steps

    def phases(plan, step, n):
        """Per-phase device durations: `n` more steps with the library's per-phase events on (CUDA events on the launch stream between
        the kernels).  They are off in the timed regions because an event between two kernels keeps the second from starting under
        programmatic dependent launch."""
        plan.set_phase_timing(True)
        for _ in range(3):
            step(stream.cuda_stream)
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for _ in range(n):
            step(stream.cuda_stream)
        p1.record(stream)
        torch.cuda.synchronize()
        ms_events = p0.elapsed_time(p1) / n
        ph = np.array([[plan.last_timing(k)[key] for key in ("sampler_ms", "integrator_ms", "scatter_ms", "total_ms")] for k in range(n)])
        launches = plan.last_timing(0)["launches"]
        plan.set_phase_timing(False)
        barrier()
        k1, k2, k3, tot = ph.mean(axis=0)
        return {"sampler_k1": float(k1), "integrator_k2": float(k2), "scatter_k3": float(k3), "sum": float(tot),
                "ms_per_step_with_phase_events": float(ms_events)}, launches

    def hbm_roofline(plan, ranges, k3_ms, tot_ms, wl, dedupe):
        n_slots = sum(e - b for b, e in ranges)
        # K3 algorithmic bytes: 16 B written per slot (A and B) + the source map it reads.  SURVEY.md 8d budgets a 4 B index per pair; the
        # packed map the kernel actually reads is smaller (fem2d_plan_source_map_info), and the smaller figure is the one used here.
        smi = plan.source_map_info()
        map_bytes = smi["map_bytes"] * (n_slots / max(plan.nnz, 1))
        alg = 16.0 * n_slots + map_bytes
        gbs = alg / (k3_ms * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": "k3_gather_kernel (DoF scatter)", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                "traffic": _ncu_traffic(wl, dedupe, world),
                "peak_source": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs, burst copy)" if peak_kind == "measured" else "fallback 6.65 TB/s",
                "algorithmic_bytes_per_launch": alg, "source_map_bytes_per_launch": map_bytes, "source_map_plain_chunks": smi["plain_chunks"],
                "survey_8d_bytes_per_launch": 20.0 * n_slots, "kernel_ms": float(k3_ms), "share_of_step": float(k3_ms / tot_ms)}

    def fp64_roofline(plan, g, k2_ms, tot_ms, wl=None, dedupe=1):
        """Integrator roofline: algorithmic FP64 lane-operations of the reference's per-pair quadrature (fem2d_plan_work_info: 8 per same-
        direction pair and point + 4 per row, 3 + 2 for cross-direction pairs; nothing fusable, tile padding and slab staging not counted)
        over the integrator's device time, against the DMUL+DADD issue rate measured in this run.  At N > 1 every rank is charged 1/N of the
        operations, so recomputed halo tiles and imbalance lower the fraction."""
        ops = plan.fp64_lane_ops(g, g) / world
        rate = ops / (k2_ms * 1e-3) / 1e9
        w = plan.work_info()
        return {"bound": "fp64_issue", "kernel": "k2_ws_kernel<4> + k2_exact_kernel<4,64> (exact per-pair integrator)", "achieved": rate, "peak": fp64_peak_gops,
                "unit": "G lane-ops/s", "frac": rate / fp64_peak_gops, "traffic": _ncu_traffic_k2(wl, dedupe, world) if wl else None,
                "peak_source": "non-fused DMUL+DADD chain measured in this run (fem2d_fp64_peak kind 1); DFMA peak is 2x and unusable: every operation of the reference order is separately rounded",
                "algorithmic_lane_ops_per_launch": ops, "same_pairs": w["same_pairs"], "cross_pairs": w["cross_pairs"],
                "kernel_ms": float(k2_ms), "share_of_step": float(k2_ms / tot_ms)}

    # ---------------------------------------------------------------------------------------------- headline: BASELINE configs[2]
    plan = F.Plan(view, device=local_rank, dedupe=bool(args.dedupe))
    nnz = plan.nnz
    ranges = rank_ranges(plan, world, rank)
    if args.emulate_world > 1:   # tuning aid: time one rank's share of an N-rank run on a single GPU
        ranges = rank_ranges(plan, args.emulate_world, args.emulate_rank)
    d_a = torch.empty(nnz, dtype=torch.float64, device=dev)
    d_b = torch.empty(nnz, dtype=torch.float64, device=dev)

    def step(st):
        plan.assemble_device_ranges(glq, d_a.data_ptr(), d_b.data_ptr(), ranges, mode=mode, stream=st)

    sampler = ClockSampler(local_rank)
    ms_step = timed(step, args.steps, max(args.warmup, 3), sampler)
    value = 2.0 * nnz / (ms_step * 1e-3)
    ph, launches_per_step = phases(plan, step, min(args.steps, 64))
    roofline = hbm_roofline(plan, ranges, ph["scatter_k3"], ph["sum"], workload, args.dedupe)
    info = plan.info

    # ------------------------------------------------------------- first-class second results: where the integrator is the step
    workloads = {}
    if not args.no_second:
        # (1) the north_star target: >= 1M-DoF anisotropically hp-refined domain, sharded like the headline at N > 1
        try:
            hp_dom = build_product_domain("hp1m")
            g = WORKLOADS["hp1m"]["glq"]
            gq = (F.gauss_quadrature_points(g), F.gauss_quadrature_points(g))
            hp = F.Plan(hp_dom.view(), device=local_rank, dedupe=True)
            hr = rank_ranges(hp, world, rank)
            ha = torch.empty(hp.nnz, dtype=torch.float64, device=dev); hb = torch.empty_like(ha)

            def hp_step(st):
                hp.assemble_device_ranges(gq, ha.data_ptr(), hb.data_ptr(), hr, mode=mode, stream=st)

            n_hp = max(3, min(args.steps, 20))
            ms_hp = timed(hp_step, n_hp, 3, tag="hp1m")
            hph, _ = phases(hp, hp_step, min(n_hp, 10))
            k2_max, tot_max = max_over_ranks(hph["integrator_k2"]), max_over_ranks(hph["sum"])
            workloads["hp1m"] = {
                "workload": WORKLOAD_TEXT["hp1m"], "n_dofs": hp.n_dofs, "nnz_upper_per_matrix": hp.nnz, "n_pairs": hp.info["n_pairs"],
                "n_blocks": hp.info["n_blocks"], "n_classes": hp.info["n_classes"], "dedupe": 1, "glq": [g, g], "steps": n_hp, "n_gpus": world,
                "ms_per_step": ms_hp, "value": 2.0 * hp.nnz / (ms_hp * 1e-3), "unit": "nnz/s", "phases_ms_rank0": hph,
                "roofline": fp64_roofline(hp, g, k2_max, tot_max, "hp1m", 1),
                "roofline_hbm": hbm_roofline(hp, hr, hph["scatter_k3"], hph["sum"], "hp1m", 1),
            }
            del hp, ha, hb, hp_dom
        except Exception as ex:  # pragma: no cover
            workloads["hp1m"] = {"error": repr(ex)}
        # (2) the headline mesh without block dedupe: every one of the 16 384 leaf blocks integrated on its own (single GPU only)
        if world == 1:
            try:
                pn = F.Plan(view, device=local_rank, dedupe=not bool(args.dedupe))

                def nd_step(st):
                    pn.assemble_device(glq, d_a.data_ptr(), d_b.data_ptr(), mode=mode, stream=st)

                ms_nd = timed(nd_step, 10, 3, tag="dedupe_off")
                nph, _ = phases(pn, nd_step, 5)
                key = f"{workload}_dedupe{int(not bool(args.dedupe))}"
                workloads[key] = {"workload": WORKLOAD_TEXT.get(workload, workload), "dedupe": int(not bool(args.dedupe)), "n_classes": pn.info["n_classes"],
                                  "ms_per_step": ms_nd, "value": 2.0 * nnz / (ms_nd * 1e-3), "unit": "nnz/s", "phases_ms": nph,
                                  "roofline": fp64_roofline(pn, w["glq"], nph["integrator_k2"], nph["sum"], workload, int(not bool(args.dedupe))),
                                  "roofline_hbm": hbm_roofline(pn, [(0, nnz)], nph["scatter_k3"], nph["sum"], workload, int(not bool(args.dedupe)))}
                del pn
            except Exception as ex:  # pragma: no cover
                workloads["dedupe_off"] = {"error": repr(ex)}

    extra = {}
    if rank == 0 and world == 1 and not args.no_extras:
        extra = run_extras(F, torch, args, domain, view, plan, glq, d_a, d_b, mode, stream, dev, local_rank, workload, nnz, hbm_peak)

    # ---- end to end through the reference-facing call: host Domain view -> host CSR arrays ------------------------------------
    if not args.no_e2e and world == 1:
        e2e = run_e2e(F, view, glq, mode, args, local_rank, rank, world, dist)

    if rank == 0 and world == 1 and not args.no_cpu and "hp1m" in workloads and "error" not in workloads["hp1m"]:
        # CPU side by side for the north_star workload too: the same full assembly by the C++ restatement, all host threads, once
        threads = os.cpu_count() or 1
        r = cpu_port_run("hp1m", threads)
        workloads["hp1m"]["cpu_baseline"] = {"value": 2.0 * r["nnz"] / r["seconds"], "unit": "nnz/s", "cores": threads, "kind": "port",
                                             "sample": f"the full workload, once ({r['dofs']} DoFs, {r['nnz']} upper entries/matrix): {r['seconds']:.2f}s = integrate "
                                                       f"{r['integrate_s']:.2f}s ({threads} threads) + serial merge {r['merge_s']:.2f}s"}
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        # the SAME workload, whole: ~30 s on 16 host cores (the serial ordered-map merge dominates, as in the reference's design)
        threads = os.cpu_count() or 1
        r = cpu_port_run(workload, threads)
        cpu_baseline = {"value": 2.0 * r["nnz"] / r["seconds"], "unit": "nnz/s", "cores": threads, "kind": "port",
                        "sample": f"the full workload, once ({r['dofs']} DoFs, {r['nnz']} upper entries/matrix): C++ restatement of the Rayon path, "
                                  f"{r['seconds']:.2f}s = integrate {r['integrate_s']:.2f}s ({threads} threads) + serial merge {r['merge_s']:.2f}s"}

    if rank == 0 and e2e and "hp1m" in e2e and "hp1m" in workloads:
        workloads["hp1m"]["e2e"] = e2e.pop("hp1m")
    if rank == 0:
        line = {
            "metric": "assembly_nnz_per_s", "value": value, "unit": "nnz/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_TEXT.get(workload, workload),
                       "mode": args.mode, "dedupe": int(args.dedupe), "n_dofs": info["n_dofs"], "nnz_upper_per_matrix": nnz, "n_pairs": info["n_pairs"],
                       "n_classes": info["n_classes"], "nnz_counted": "2 x nnz_upper (A and B)",
                       "launch": launch_kind,
                       "l2_policy": "no flush: each step streams > 0.92 GB (A/B value arrays + source map) >> 126 MB L2",
                       "parallelism": f"row blocks x{world} (Elem-type rows + edge-type rows per rank), no collective" if world > 1 else "single GPU"},
            "phases_ms": dict(ph, note="extra steps with per-phase events on, right after the timed region"),
            "roofline": roofline,
            "roofline_fp64": workloads.get("hp1m", {}).get("roofline"),
            "workloads": workloads,
            "fp64_peak": fp64,
            "cpu_baseline": cpu_baseline,
            "e2e": e2e,
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": sampler.summary(),
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_extras(F, torch, args, domain, view, plan, glq, d_a, d_b, mode, stream, dev, local_rank, workload, nnz, hbm_peak):
    """Context series (N = 1 only): the other BASELINE configs that fit one GPU, Q-scaling, the second basis space, xy_fields, and the
    re-ordered integrator modes with their own HBM roofline and both tolerance verdicts."""
    import numpy as np
    extra = {}

    def time_plan(pl, gq, a, b, reps=20, **kw):
        for _ in range(3):
            pl.assemble_device(gq, a.data_ptr(), b.data_ptr(), stream=stream.cuda_stream, **kw)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for _ in range(reps):
            pl.assemble_device(gq, a.data_ptr(), b.data_ptr(), stream=stream.cuda_stream, **kw)
        g1.record(stream)
        torch.cuda.synchronize()
        return g0.elapsed_time(g1) / reps

    others = {}
    for wl in ("cfg2", "cfg4"):
        try:
            dom = build_product_domain(wl)
            g = WORKLOADS[wl]["glq"]
            gq = (F.gauss_quadrature_points(g), F.gauss_quadrature_points(g))
            pl = F.Plan(dom.view(), device=local_rank, dedupe=bool(args.dedupe))
            ta = torch.empty(pl.nnz, dtype=torch.float64, device=dev); tb = torch.empty_like(ta)
            ms = time_plan(pl, gq, ta, tb, mode=mode)
            others[wl] = {"n_dofs": pl.n_dofs, "nnz_upper_per_matrix": pl.nnz, "n_classes": pl.info["n_classes"], "glq": [g, g], "ms_per_step": ms,
                          "value": 2.0 * pl.nnz / (ms * 1e-3), "unit": "nnz/s"}
            if wl == "cfg4":
                ms2 = time_plan(pl, gq, ta, tb, mode=mode, basis=F.HierMaxOrtho)
                others["cfg4_hier_max_ortho"] = {"glq": [g, g], "ms_per_step": ms2, "value": 2.0 * pl.nnz / (ms2 * 1e-3), "unit": "nnz/s"}
            del pl, ta, tb
        except Exception as ex:  # pragma: no cover
            others[wl] = {"error": str(ex)}
    extra["other_configs"] = others
    series = {}
    try:
        # Q-scaling: the headline mesh with the reference's default quadrature (`glq_grid_dim = None` -> default_ngq(order)
        # points per side, basis.rs:172-177)
        ng = int(F.default_ngq(max(domain.mesh.max_expansion_orders())))
        gq = (F.gauss_quadrature_points(ng), F.gauss_quadrature_points(ng))
        ms = time_plan(plan, gq, d_a, d_b, mode=mode)
        series[workload + "_glq_default"] = {"glq": [ng, ng], "ms_per_step": ms, "value": 2.0 * nnz / (ms * 1e-3), "unit": "nnz/s"}
        # UniformFieldSpace::xy_fields (fields.rs:63-127) on the headline mesh, [16,16] points per leaf Elem,
        # host solution vector in, host field arrays out (wall clock around the C-ABI call)
        import ctypes as _C
        dens, cap = 16, domain.mesh.num_elems
        sol = np.cos(np.arange(domain.num_dofs) * 0.37) + 0.25
        ids = np.zeros(cap, dtype=np.uint32)
        fx = np.zeros((cap, dens, dens)); fy = np.zeros((cap, dens, dens))
        n_out = _C.c_uint64()
        ptr = lambda arr, t: arr.ctypes.data_as(_C.POINTER(t))

        def fields_call():
            st = F._L.fem2d_xy_fields(_C.byref(view.c), int(local_rank), F.HierPoly.kind, _C.c_uint32(dens), ptr(sol, _C.c_double), _C.c_uint64(cap),
                                      _C.byref(n_out), ptr(ids, _C.c_uint32), ptr(fx, _C.c_double), ptr(fy, _C.c_double))
            if st != 0:
                raise RuntimeError(F._L.fem2d_last_error().decode())

        fields_call()
        t0 = time.perf_counter()
        for _ in range(3):
            fields_call()
        ms = (time.perf_counter() - t0) / 3 * 1e3
        series[workload + "_xy_fields_16x16"] = {"leaf_elems": int(n_out.value), "ms_per_call_host_to_host": ms,
                                                 "points_per_s": n_out.value * dens * dens / (ms * 1e-3),
                                                 "note": "wall clock around fem2d_xy_fields: pageable host solution in, pageable host field arrays out"}
    except Exception as ex:  # pragma: no cover
        series["error"] = str(ex)
    extra["other_series"] = series
    # re-ordered integrator modes (opt-in, never the default): time, HBM roofline of the whole step, and both tolerance verdicts against
    # the EXACT (reference-order) values of the same plan
    fast = {}
    try:
        pn = F.Plan(view, device=local_rank, dedupe=False)
        ref_a = torch.empty(nnz, dtype=torch.float64, device=dev); ref_b = torch.empty_like(ref_a)
        pn.assemble_device(glq, ref_a.data_ptr(), ref_b.data_ptr(), mode=F.MODE_EXACT, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        for name, md in (("sumfact", F.MODE_SUMFACT), ("dmma", F.MODE_DMMA)):
            ms = time_plan(pn, glq, d_a, d_b, reps=10, mode=md)
            pn.set_phase_timing(True)
            pn.assemble_device(glq, d_a.data_ptr(), d_b.data_ptr(), mode=md, stream=stream.cuda_stream)
            torch.cuda.synchronize()
            t = pn.last_timing(0)
            pn.set_phase_timing(False)
            verdict = {}
            for nm, got, ref in (("A", d_a, ref_a), ("B", d_b, ref_b)):
                err = (got - ref).abs()
                scale = float(ref.abs().max())
                lit = int((err > torch.clamp(1e-12 * ref.abs(), min=1e-14)).sum())
                sca = int((err > torch.clamp(1e-12 * ref.abs(), min=1e-14 * scale)).sum())
                verdict[nm] = {"violations_literal_1e-12_rel_1e-14_abs": lit, "violations_scale_aware_floor_1e-14_x_max": sca, "max_abs": scale,
                               "max_abs_err": float(err.max())}
            # bytes the step must move: V written by the integrator (16 B per integrated pair, dedupe off) and read by the scatter (16 B per slot)
            # + A, B out (16 B per slot) + the packed source map
            step_bytes = 16.0 * pn.info["n_pairs"] + 32.0 * nnz + pn.source_map_info()["map_bytes"]
            fast[name] = {"dedupe": 0, "ms_per_step": ms, "value": 2.0 * nnz / (ms * 1e-3), "unit": "nnz/s", "integrator_ms": t["integrator_ms"],
                          "scatter_ms": t["scatter_ms"],
                          "roofline": {"bound": "hbm", "kernel": f"{name} integrator + scatter (whole step)", "achieved": step_bytes / (ms * 1e-3) / 1e9,
                                       "peak": hbm_peak, "unit": "GB/s", "frac": step_bytes / (ms * 1e-3) / 1e9 / hbm_peak, "traffic": None,
                                       "algorithmic_bytes_per_step": step_bytes},
                          "tolerance_vs_exact_mode": verdict}
        del pn, ref_a, ref_b
    except Exception as ex:  # pragma: no cover
        fast["error"] = repr(ex)
    extra["reordered_modes"] = fast
    return extra


def rank_ranges(plan, world, rank):
    """Slot ranges owned by `rank`: everything for one rank; otherwise its block of the single-Elem (Elem-type) rows and its block of
    the shared (edge-type) rows (fem2d_plan_row_blocks_split)."""
    if world == 1:
        return [(0, plan.nnz)]
    b1, b2 = plan.row_blocks_split(world)
    return [(int(b1[rank]), int(b1[rank + 1])), (int(b2[rank]), int(b2[rank + 1]))]


def run_e2e(F, view, glq, mode, args, local_rank, rank, world, dist):
    """The reference-facing call with HOST buffers, every step: fem2d_galerkin_sample_gep_hcurl_multi = host planner + symbolic phase on
    every device + K1/K2/K3 + D2H of A, B and the compressed pattern + host expansion of rows[] / cols[], ONE process driving all N GPUs
    (what a single Rust caller of the drop-in gets).  Under torchrun rank 0 makes the call on devices 0..N-1 while the other ranks wait at
    a barrier.  Wall clock around the synchronous C-ABI call.  Variants at N = 1: pageable buffers exactly as INTEGRATION.md's first
    listing allocates them (plain vectors), and the ordered-map rebuild a Rust caller pays afterwards (std::map stand-in)."""
    import ctypes as C
    import numpy as np
    import torch

    names = ["elem_element", "elem_parent", "elem_loc", "element_p0", "element_p3", "element_eps_re", "element_mu_re", "bs_off", "bs_i", "bs_j",
             "bs_dir", "bs_dof"]
    steps = max(1, min(args.steps, args.e2e_steps))
    out = None
    if rank == 0:
        # size of the result: a first call with capacity 0 reports nnz and fails with BAD_ARGUMENT
        probe0 = F.Plan(view, device=0, dedupe=True)
        nnz = probe0.nnz
        del probe0
        # inputs: the flattened Domain arrays in pinned host memory
        pinned = {}
        h2d = 0
        cv = type(view.c)()
        C.memmove(C.byref(cv), C.byref(view.c), C.sizeof(cv))
        for nm in names:
            src = np.ascontiguousarray(getattr(view, nm))
            t = torch.from_numpy(src.copy()).pin_memory()
            pinned[nm] = t
            h2d += t.numel() * t.element_size()
            fld = dict(type(cv)._fields_)[nm]
            setattr(cv, nm, C.cast(t.data_ptr(), fld))
        pview = F.DomainView(cv, keepalive=pinned)
        devices = list(range(world))
        h_rows = torch.empty(nnz, dtype=torch.int32).pin_memory()
        h_cols = torch.empty(nnz, dtype=torch.int32).pin_memory()
        h_a = torch.empty(nnz, dtype=torch.float64).pin_memory()
        h_b = torch.empty(nnz, dtype=torch.float64).pin_memory()
        ptrs = (h_rows.data_ptr(), h_cols.data_ptr(), h_a.data_ptr(), h_b.data_ptr(), nnz)

        def one(v=pview, p=ptrs):
            got = F.galerkin_sample_gep_hcurl_multi(v, glq, devices, mode=mode, out=p)
            assert got == nnz

        for _ in range(4):
            one()             # (the call is synchronous: every device is idle when it returns)
        per_step = []
        t0 = time.perf_counter()
        for _ in range(steps):
            t1 = time.perf_counter()
            one()
            per_step.append(round(1e3 * (time.perf_counter() - t1), 2))
        sec = time.perf_counter() - t0
        tm = (C.c_double * (4 + 4 * world))()
        F._L.fem2d_debug_multi_timing(tm, C.c_uint32(4 + 4 * world))
        breakdown = {"host_planner_ms": round(tm[0], 2), "call_ms": round(tm[1], 2),
                     "per_device_ms": [{"symbolic": round(tm[4 + 4 * r], 2), "row_block_split": round(tm[5 + 4 * r], 2), "numeric_d2h_expand": round(tm[6 + 4 * r], 2),
                                        "plan_release": round(tm[7 + 4 * r], 2)} for r in range(world)]}
        # what crosses PCIe device -> host per step: A, B and the compressed pattern (CSR row offsets + column runs; rows[] / cols[] are expanded on host)
        probe = F.Plan(pview, device=local_rank, dedupe=True)
        xfer = probe.pattern_transfer_info()
        t_sym = {"symbolic_host_ms": round(probe.info["symbolic_host_us"] / 1e3, 2), "symbolic_device_ms": round(probe.info["symbolic_device_us"] / 1e3, 2)}
        del probe
        d2h = nnz * 16 + (xfer["row_offset_bytes"] + xfer["col_run_bytes"])
        out = {"value": 2.0 * nnz * steps / sec, "unit": "nnz/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * sec / steps,
               "steps": steps, "n_devices": world, "call": "fem2d_galerkin_sample_gep_hcurl_multi (one process, one host thread per device, pinned host buffers)",
               "includes": "host planner + symbolic phase (pattern + source map) on every device + K1/K2/K3 + D2H of A, B and the compressed pattern (CSR row offsets, column runs) into pinned host buffers + host expansion of rows[] and cols[]",
               "per_step_ms": per_step, "last_call_breakdown": breakdown, "symbolic_phase_of_one_device": t_sym}
        if world == 1 and not args.no_extras:
            # the buffers a caller following INTEGRATION.md's plain-vector listing has: pageable view arrays, pageable outputs
            n_rows = np.empty(nnz, dtype=np.uint32); n_cols = np.empty(nnz, dtype=np.uint32); n_a = np.empty(nnz); n_b = np.empty(nnz)
            n_rows.fill(0); n_cols.fill(0); n_a.fill(0.0); n_b.fill(0.0)        # touch the pages once, as a Vec::with_capacity + resize would
            pp = (n_rows.ctypes.data, n_cols.ctypes.data, n_a.ctypes.data, n_b.ctypes.data, nnz)
            one(view, pp)
            t0 = time.perf_counter()
            reps = min(3, steps)
            for _ in range(reps):
                one(view, pp)
            pg_ms = 1e3 * (time.perf_counter() - t0) / reps
            assert np.array_equal(n_a.view(np.uint64), h_a.numpy().view(np.uint64)) and np.array_equal(n_rows, h_rows.numpy().view(np.uint32))
            reb = F._L.fem2dh_ordered_map_rebuild_seconds(C.c_uint64(nnz), n_rows.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                          n_cols.ctypes.data_as(C.POINTER(C.c_uint32)), n_a.ctypes.data_as(C.POINTER(C.c_double)))
            out["pageable_buffers"] = {"ms_per_step": pg_ms, "value": 2.0 * nnz / (pg_ms * 1e-3), "unit": "nnz/s",
                                       "note": "same call, pageable view arrays and pageable outputs (plain vectors, INTEGRATION.md first listing); pinned outputs from fem2d_host_alloc are the documented path"}
            out["caller_side_rebuild"] = {"seconds_per_matrix": reb, "what": "std::map<[u32;2], f64> filled from the sorted arrays with end hints + its destruction: stand-in for SparseMatrix::from_sorted_upper_tri (BTreeMap bulk build) in the Rust shim; NOT inside e2e.value",
                                          "e2e_ms_including_two_rebuilds_pinned": 1e3 * sec / steps + 2e3 * reb}
    if rank == 0 and not args.no_second:
        # the north_star statement end to end: ONE call on a host view of the >= 1M-DoF hp-mesh -> host CSR arrays, all N GPUs
        try:
            hp_view = build_product_domain("hp1m").view()
            g = WORKLOADS["hp1m"]["glq"]
            gq = (F.gauss_quadrature_points(g), F.gauss_quadrature_points(g))
            pr = F.Plan(hp_view, device=0, dedupe=True)
            n_hp = pr.nnz
            t_host = pr.info["symbolic_host_us"] / 1e3
            del pr
            hb = [torch.empty(n_hp, dtype=t).pin_memory() for t in (torch.int32, torch.int32, torch.float64, torch.float64)]
            hp_ptrs = (hb[0].data_ptr(), hb[1].data_ptr(), hb[2].data_ptr(), hb[3].data_ptr(), n_hp)
            F.galerkin_sample_gep_hcurl_multi(hp_view, gq, devices, mode=mode, out=hp_ptrs)
            ts = []
            for _ in range(3):
                t1 = time.perf_counter()
                F.galerkin_sample_gep_hcurl_multi(hp_view, gq, devices, mode=mode, out=hp_ptrs)
                ts.append(1e3 * (time.perf_counter() - t1))
            out["hp1m"] = {"ms_per_call": sum(ts) / len(ts), "per_call_ms": [round(x, 1) for x in ts], "value": 2.0 * n_hp / (sum(ts) / len(ts) * 1e-3), "unit": "nnz/s",
                           "n_devices": world, "host_planner_ms": round(t_host, 1), "d2h_bytes": int(n_hp * 16),
                           "note": "fem2d_galerkin_sample_gep_hcurl_multi on the 1,380,549-DoF hp-mesh, host view in, pinned host rows/cols/A/B out; the serial host planner "
                                   "(class hashing of 71,086 blocks, work items, packs) is the largest part"}
            del hb
            F.trim_cache()
        except Exception as ex:  # pragma: no cover
            out["hp1m"] = {"error": repr(ex)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="exact", choices=["exact", "sumfact", "dmma"])
    ap.add_argument("--dedupe", type=int, default=1)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every step directly instead of replaying a CUDA graph of it")
    ap.add_argument("--no-second", action="store_true", help="skip the hp1m / dedupe-off first-class second results")
    ap.add_argument("--emulate-world", type=int, default=1)
    ap.add_argument("--emulate-rank", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
