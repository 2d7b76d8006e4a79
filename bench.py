#!/usr/bin/env python
"""bench.py — candidate configs/sec of the Analyze -> Optimize hot path (BASELINE.json metric).

One "step" = one reconcile pass over the workload: system upload, Server.Calculate for every
(server, accelerator) pair (wva_analyze_pairs), the (server x accelerator x replicas x batch)
candidate sweep with per-server argmin (wva_analyze_grid), the assignment (wva_solve) and the
per-type totals (wva_allocate_by_type, + one NCCL all-gather and a rank-order sum when N > 1).

  value : whole-job candidates/s with the system image already resident in HBM (device timed,
          CUDA events on the library's stream, max over ranks)
  e2e   : same metric through the public C-ABI with HOST buffers: H2D of the image and D2H of
          the decisions inside the timed region
  roofline / fp64 : the sweep kernel against the measured HBM peak (as BASELINE.json asks) and
          against the FP64 issue rate that actually bounds it
  cpu_baseline : the oracle (CPU restatement of the Go path) on a bounded sample, same box

`--impl reference` times the reference's CPU implementation (the oracle port: no Go toolchain
exists here) on the host cores instead.
"""
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

METRIC = "candidate_configs_per_sec"
UNIT = "candidates/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", d
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)", {}


def ncu_traffic(kernel, cfg_id):
    """dram__bytes_read.sum + dram__bytes_write.sum of `kernel` from the committed `ncu --set full` summary
    of the same workload (profiles/ncu_full_r01_cfg<k>.json), per launch; None when there is none."""
    p = os.path.join(ROOT, "profiles", "ncu_full_r01_cfg%d.json" % cfg_id)
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        for k in json.load(open(p)):
            if k["kernel"].split("(")[0] == kernel:
                tot = 0.0
                for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    v, u = k[key].split()
                    tot += float(v.replace(",", "")) * mult[u]
                return tot
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def workload(cfg_id, world, servers_per_rank=None):
    """System image of BASELINE config `cfg_id`; weak scaling: every rank owns a full copy of the
    config's server count, the job has world x that many servers."""
    import wva_import
    wva = wva_import.load()
    c = dict(wva.synth.CONFIGS[cfg_id])
    if cfg_id == 1:
        img = wva.synth.config1()
        per_rank = 1
        if world > 1:
            raise SystemExit("config 1 is a single-server case; use --config 2+ for multi-GPU")
    else:
        per_rank = servers_per_rank or c["S"]
        img = wva.synth.make_system(per_rank * world, c["A"], seed=cfg_id, n_types=c["T"])
    return wva, img, c, per_rank


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores (oracle port;
    the Go toolchain is absent so the reference itself cannot run, DESIGN.md "Oracle")."""
    if rank != 0:
        return
    import oracle
    wva, img, c, per_rank = workload(args.config, 1)
    threads = oracle.hardware_threads()
    # bounded sample: a slice of servers sized for a few seconds per step
    n_srv = max(1, min(img.S, args.ref_servers))
    cand = n_srv * img.A * c["r_max"] * c["b_max"]

    def step():
        sub = img.shard(0, n_srv)
        pairs, feas, _ = oracle.analyze_pairs(sub, threads=threads)
        oracle.analyze_grid(sub, c["r_max"], c["b_max"], want_cube=False, threads=threads)
        acc, chosen = oracle.solve(sub, pairs, feas, unlimited=True)
        oracle.allocate_by_type(sub, acc, chosen)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = cand / dt
    sample = "%d of %d servers of config %d (%d candidates + %d pairs per step)" % (n_srv, img.S, args.config, cand, n_srv * img.A)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 chain / f32 metrics", "data": "synthetic",
            "config": {"workload": "BASELINE config %d: %d servers x %d accel x r<=%d x b<=%d" %
                       (args.config, img.S, img.A, c["r_max"], c["b_max"]), "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


_REAL_STDOUT = None


def _quiet_stdout():
    """Keep stdout for the one JSON line: anything a library prints there (NCCL's version banner, ...)
    goes to stderr instead."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json config index (1-based): 2 = 32 models x 4 accel")
    ap.add_argument("--servers-per-rank", type=int, default=None)
    ap.add_argument("--ref-servers", type=int, default=4, help="servers per step of the CPU arms' bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--limited", action="store_true",
                    help="capacity-constrained assignment (SolveGreedy, PriorityExhaustive): capacities = 60 %% of the "
                         "unconstrained demand; multi-GPU ranks gather the candidate rows and solve redundantly")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build_cuda()
    wva, img, c, per_rank = workload(args.config, world, args.servers_per_rank)
    from inferno_autoscaler_b200 import binding
    abi = wva.abi

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = binding.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local_rank))
    R, B = c["r_max"], c["b_max"]
    first = rank * per_rank
    cand_rank = per_rank * img.A * R * B
    cand_total = cand_rank * world
    l2_flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    # the metric cube (33 B per candidate) is materialised in HBM when it fits comfortably
    want_cube = cand_rank * 33 <= 24 * (1 << 30)
    rows_mode = per_rank * img.A * R >= 32768

    from inferno_autoscaler_b200 import distributed as D
    dev = torch.device("cuda", local_rank)
    if args.limited:
        # capacities that bind: 60 % of the demand of an unconstrained solve, computed once, untimed
        ctx.upload(img)
        ctx.analyze_pairs(download=False)
        acc0, ch0 = ctx.solve(unlimited=True)
        wva.synth.set_capacity_from_demand(img, ch0.acc, ch0.num_replicas, fraction=0.6)

    totals_t = None

    def solve_step(download):
        if not args.limited:
            return ctx.solve(unlimited=True, download=download)
        if world > 1:
            with torch.cuda.stream(stream):
                D.gather_pair_rows_device(ctx, img.S, img.A, world, dev)
        return ctx.solve(unlimited=False, policy=abi.POLICY_PRIORITY_EXHAUSTIVE, download=download)

    def allreduce_totals():
        """the one collective of the path: all-gather of the per-type {count, cost} partials (NCCL) + rank-order sum."""
        nonlocal totals_t
        if world == 1:
            return
        if totals_t is None:
            totals_t = D.TotalsExchange(ctx, img.T, dev, stream)
        totals_t(sync=False)

    def step_device():
        """hot path with the image resident in HBM; decisions stay in HBM."""
        ctx.analyze(R, B, want_cube=want_cube)       # Server.Calculate for all pairs || candidate sweep
        solve_step(False)
        ctx.allocate_by_type(download=False)
        allreduce_totals()

    def step_e2e():
        """public API with host buffers: H2D image, D2H decisions."""
        ctx.upload(img)
        ctx.set_shard(first, per_rank)
        ctx.analyze(R, B, want_cube=False)
        pairs = ctx.pairs_fetch()
        best = ctx.grid_fetch()
        chosen = solve_step(True)
        tot = ctx.allocate_by_type()
        allreduce_totals()
        return pairs, best, chosen, tot

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx.upload(img)
    ctx.set_shard(first, per_rank)
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    grid_kernel_us, phase_us = [], {k: [] for k in ("pairs", "grid", "solve", "totals", "grid_light_kernel", "grid_heavy")}
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        l2_flush.fill_(i & 0xff)                      # evict L2 between timed iterations (not timed)
        barrier()                                     # ranks start every timed step together (not timed)
        ev[i][0].record(stream)
        step_device()
        ev[i][1].record(stream)
        torch.cuda.synchronize()
        grid_kernel_us.append(ctx.phase_usec(abi.PHASE_GRID_KERNEL) + ctx.phase_usec(abi.PHASE_GRID_HEAVY))
        for k, ph in (("pairs", abi.PHASE_PAIRS), ("grid", abi.PHASE_GRID), ("solve", abi.PHASE_SOLVE), ("totals", abi.PHASE_TOTALS),
                      ("grid_light_kernel", abi.PHASE_GRID_KERNEL), ("grid_heavy", abi.PHASE_GRID_HEAVY)):
            phase_us[k].append(ctx.phase_usec(ph))
    barrier()
    wall = time.perf_counter() - t_wall0
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    counters = ctx.grid_counters()
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = cand_total / (ms_per_step * 1e-3)

    # ---- e2e through the public API ------------------------------------------------------
    for _ in range(2):
        out = step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    pairs, best, chosen, tot = out
    h2d = img.nbytes()
    d2h = (sum(getattr(pairs[0], n).nbytes for n, _ in abi.ALLOC_FIELDS) + pairs[1].nbytes) * per_rank // img.S
    d2h += best.nbytes + (sum(getattr(chosen[1], n).nbytes for n, _ in abi.ALLOC_FIELDS) + chosen[0].nbytes) * per_rank // img.S
    d2h += tot[0].nbytes + tot[1].nbytes

    if rank == 0:
        hbm_peak, peak_src, peaks = load_peaks()
        k_us = float(np.mean(grid_kernel_us))
        bytes_per_cand = 33.0 if want_cube else 0.0               # 32 B AnalysisMetrics + 1 status byte when the cube is materialised
        alg_bytes = bytes_per_cand * cand_rank + 88.0 * per_rank * img.A + 32.0 * per_rank
        achieved = alg_bytes / (k_us * 1e-6) / 1e9
        # FP64 view (estimate; the measured figure is ncu's sm__inst_executed_pipe_fp64 in profiles/): a chain-state
        # update is ~8 FP64-pipe instructions, a certified closed-form tail (exp, log1p, 5 divisions) ~150
        fp64_ops = counters["steps_executed"] * 8.0 + counters["candidates_ok"] * 150.0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 chain / f32 metrics", "data": "synthetic",
            "config": {"workload": "BASELINE config %d: %d servers/GPU x %d accel x replicas 1-%d x batch 1-%d (+ %d Server.Calculate pairs, unlimited solve, per-type totals)"
                       % (args.config, per_rank, img.A, R, B, per_rank * img.A),
                       "candidates_per_step": cand_total, "pairs_per_step": per_rank * img.A * world,
                       "l2": "flushed between timed iterations (256 MB write)",
                       "cube": "materialised in HBM (33 B/candidate)" if want_cube else "not materialised (winners only)",
                       "assignment": "greedy, capacity caps at 60% of demand, PriorityExhaustive" if args.limited else "unlimited"},
            "clocks": clocks,
            "e2e": {"value": cand_total / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_s * 1e3},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": ncu_traffic("k_grid_rows" if rows_mode else "k_grid", args.config) if want_cube else None, "peak_source": peak_src, "kernel": ("k_grid_rows" if rows_mode else "k_grid") + " (+ k_grid_list for uncertified chains)", "kernel_us": k_us,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "the sweep is FP64/issue bound, not HBM bound (see fp64); HBM fraction reported because BASELINE.json asks for it"},
            "fp64": {"chain_steps_executed": counters["steps_executed"], "chain_steps_reference": counters["steps_algorithmic"],
                     "truncation_ratio": counters["steps_algorithmic"] / max(1, counters["steps_executed"]),
                     "fp64_inst_per_s": fp64_ops / (k_us * 1e-6), "candidates_analysed": counters["candidates_ok"],
                     "lists": ctx.grid_list_sizes()},
            "phases_ms": {k: float(np.mean(v)) / 1e3 for k, v in phase_us.items()},
            "wall_s_timed_region": wall,
        }
        if not args.no_cpu_baseline and world == 1:
            import oracle
            threads = oracle.hardware_threads()
            n_srv = max(1, min(img.S, args.ref_servers))
            sub = img.shard(0, n_srv)
            t0 = time.perf_counter()
            p_, f_, _ = oracle.analyze_pairs(sub, threads=threads)
            oracle.analyze_grid(sub, R, B, want_cube=False, threads=threads)
            a_, c_ = oracle.solve(sub, p_, f_, unlimited=True)
            oracle.allocate_by_type(sub, a_, c_)
            dt = time.perf_counter() - t0
            cand = n_srv * img.A * R * B
            line["cpu_baseline"] = {"value": cand / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d of %d servers (%d candidates + %d pairs), %.1f s" % (n_srv, per_rank, cand, n_srv * img.A, dt)}
        _emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
