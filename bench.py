#!/usr/bin/env python
"""bench.py -- objective evaluations per second of a gadget3 model on N B200s.

One "step" = one pass of the hot path over one batch of synthetic parameter vectors per GPU
(BASELINE.json configs[1]: multiple-substocks, 65,536 vectors, nll only). Prints ONE JSON line.

  value      evals/s with the parameter batches already resident in HBM (device pointers through
             g3cuda_eval_batch_device), CUDA events on the launching stream, max over ranks
  e2e        evals/s through the reference-facing call (g3_cuda_adfun(...).fn_batch -> g3cuda_eval_batch)
             with HOST buffers: pinned host -> device copy of the parameters and device -> host read of
             the nll vector inside the timed region
  roofline   fp64-pipe roofline of the evaluation kernel: algorithmic flop per evaluation (SURVEY.md
             section 8d) x evaluations per launch / mean launch duration, against the DFMA
             throughput measured live on this GPU (MEASURED_PEAKS.json has no fp64 entry); `traffic` is the
             DRAM traffic of the same kernel from this round's ncu capture (profiles/r02_traffic.json)
  parity_spot_check  the timed path's results against the CPU checker on a slice of the timed batch (untimed);
             the run FAILS if they are outside the contract (1e-10 on nll and components, plus the stated
             conditioning of g3l_understocking: oracle/parity.py)
  cpu_baseline  the CPU restatement of the reference objective (oracle/) on the box's host cores, bounded sample
  extra_configs  short timed runs of the other BASELINE.json configs (same fields); at N > 1 also config 5 as
             specified: a 1000 x 1000 likelihood-profile grid strong-scaled through gadget3_b200.shard.eval_sharded
             with the NCCL gather inside e2e

`--impl reference` times that CPU implementation alone (all host threads) and prints the same line.
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
sys.path.insert(0, ROOT)

METRIC = "objective_evals_per_sec"
UNIT = "evals/s"
# Algorithmic flop per evaluation, reference semantics (DESIGN.md "flop model")
ALG_FLOP = {"C1": 4.39e6, "C2": 2.60e7, "C3": 2.79e9, "C4": 1.31e8, "C5": 4.9e7}
WORKLOAD = {"C1": "introduction-single-stock, nll only", "C2": "multiple-substocks imm/mat, nll only",
            "C3": "multiple-substocks imm/mat, nll + forward-mode gradient over all 71 optimised parameters",
            "C4": "ling-style: 2 stocks x 35 length groups, 3 fleets, constant maturity, 40 years quarterly, nll only",
            "C5": "4-area predator-prey with migration and a stock predator, 20 years quarterly, nll only"}
MODEL_OF = {"C1": "C1", "C2": "C2", "C3": "C2", "C4": "C4", "C5": "C5"}
DEFAULT_BATCH = {"C1": 65536, "C2": 65536, "C3": 8192, "C4": 65536, "C5": 131072}
EXTRA_STEPS = {"C1": 3, "C3": 2, "C4": 2, "C5": 2}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(ALG_FLOP))
    ap.add_argument("--batch", type=int, default=0, help="parameter vectors per GPU per step (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs of the other configs")
    ap.add_argument("--kernel", type=int, default=0, help="0 = automatic dispatch, 3 = thread kernel, 4 = tile kernel (profiling)")
    ap.add_argument("--tile-warps", type=int, default=0, help="warps per tile of the tile kernel (0 = planner's choice)")
    ap.add_argument("--tile-seg", type=int, default=0, help="rows per unit of the tile kernel (0 = planner's choice, -1 = whole columns)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [ln.split(", ") for ts, ln in self.lines if t0 <= ts <= t1 + 0.2] or [ln.split(", ") for _, ln in self.lines]
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                "samples": len(sm)}


def workload_cfg(config, batch, n_par, n_opt, n_steps, world):
    """The `config` object of a line: the same dict from the CUDA arm and from `--impl reference`."""
    return {"workload": f"{config}: {WORKLOAD[config]}, {batch} parameter vectors per GPU per step",
            "n_par": int(n_par), "n_optimised": int(n_opt), "batch_per_gpu": int(batch),
            "steps_per_eval": int(n_steps), "parallelism": f"batch-sharded x{world}"}


def load_oracle(model, params):
    from oracle import oracle_lib as ol
    try:
        lib = ol.load(ol.build(native=True))          # -O3 -march=native, like TMB's default flags
    except Exception:
        lib = ol.load()
    return ol.Oracle(*model.pack(params), lib=lib)


def cpu_reference(orc, model, params, config, batch_fn, seconds_target=12.0, tix=None):
    """Time the CPU implementation of the path on all host cores over a bounded sample."""
    cores = os.cpu_count() or 1
    probe = batch_fn(2 * cores, 99)
    t = time.perf_counter(); orc.eval_batch(probe, tangent_idx=tix, nthreads=cores); per = (time.perf_counter() - t) / len(probe)
    n = int(min(max(seconds_target / max(per, 1e-9), 2 * cores), 1 << 16))
    n = max(cores, n // cores * cores)
    sample = batch_fn(n, 7)
    t = time.perf_counter(); nll, _, _ = orc.eval_batch(sample, tangent_idx=tix, nthreads=cores); dt = time.perf_counter() - t
    assert np.isfinite(nll).all()
    out = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"{n} vectors of the {config} batch, {cores} threads, {dt:.1f} s wall"}
    if tix is not None:
        # The port differentiates in FORWARD mode (dual numbers, 8 directions a pass). The reference's g3_tmb_adfun uses
        # TMBad REVERSE mode (R/run_tmb.R:1126-1127), whose gradient typically costs 3-5 values whatever the number of
        # parameters: the fair CPU figure for (value + gradient) is about a quarter of the value-only rate.
        nv = max(cores, n * 4)
        vs = batch_fn(nv, 8)
        t = time.perf_counter(); orc.eval_batch(vs, nthreads=cores); dv = time.perf_counter() - t
        vrate = nv / dv
        out["reverse_mode_estimate"] = {"value": vrate / 4.0, "unit": "(value+gradient)/s",
                                        "how": f"value-only rate of the port on the same {cores} threads ({vrate:.0f} evals/s) / 4, "
                                               "the cost of a TMBad reverse sweep: compare the GPU figure with THIS, not with the forward-mode port"}
    else:
        # for information: the reference's own generated objective (baseline/*.cpp compiled against oracle/tmb_shim,
        # prebuilt in oracle/_ref). Its eager containers make it slower than the port, so the port stays the baseline.
        try:
            from concurrent.futures import ThreadPoolExecutor
            from oracle import ref_lib
            name = MODEL_OF.get(config, config)
            if ref_lib.available(name):
                refs = [ref_lib.RefModel(name, model, params) for _ in range(cores)]
                nref = 4 * cores
                rows = batch_fn(nref, 11)
                t = time.perf_counter()
                with ThreadPoolExecutor(cores) as ex:
                    list(ex.map(lambda k: refs[k].eval_batch(rows[k::cores]), range(cores)))
                dtr = time.perf_counter() - t
                out["reference_generated_code"] = {"value": nref / dtr, "unit": UNIT, "cores": cores,
                                                   "what": "the reference's checked-in generated C++ under oracle/tmb_shim"}
        except Exception as e:        # the prebuilt library is optional
            out["reference_generated_code"] = {"unavailable": str(e)[:120]}
    return out


def traffic_of(config):
    """DRAM bytes per evaluation of the evaluation kernel from this round's ncu capture (None if none was made)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        e = t.get(config)
        return (e["dram_bytes_per_eval"], e["capture"]) if e else (None, None)
    except OSError:
        return None, None


def measure(config, batch, steps, warmup, env, a, main_run):
    """Time one config on this rank's GPU; returns the fields of a bench line (max over ranks already taken)."""
    import torch
    from gadget3_b200.configs import CONFIGS, parameter_batch
    from gadget3_b200.run_cuda import g3_cuda_adfun
    rank, world, local, dev, dist = env["rank"], env["world"], env["local"], env["dev"], env["dist"]
    want_grad = config == "C3"
    model, params = CONFIGS[MODEL_OF[config]]()
    opt = params.optimised_indices()
    base = params.values()

    def batch_fn(n, seed):
        full = np.tile(base, (n, 1))
        full[:, opt] = parameter_batch(params, n, seed=seed)
        return full

    fn = g3_cuda_adfun(model, params, device=local)
    h = fn.handle
    if a.kernel:
        h.set_kernel(a.kernel)
    if main_run and (a.tile_warps or a.tile_seg):
        h.set_tile_partition(a.tile_warps, a.tile_seg)
    B, P = batch, h.n_par

    # ---- inputs: NB distinct batches so that the parameter stream (NB x B x P x 8 B) exceeds the 126 MB L2
    nb = max(2, int(np.ceil(140e6 / (B * P * 8))))
    host = [torch.from_numpy(batch_fn(B, 1000 + 17 * rank + i)).pin_memory() for i in range(nb)]
    d_par = [x.to(dev, non_blocking=True) for x in host]
    d_nll = torch.empty(B, dtype=torch.float64, device=dev)
    Kt = len(opt) if want_grad else 0
    tix = opt.astype(np.int32) if want_grad else None
    d_tidx = torch.from_numpy(tix).to(dev) if want_grad else None
    d_grad = torch.empty((B, Kt), dtype=torch.float64, device=dev) if want_grad else None
    stream = torch.cuda.current_stream(dev)
    torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device(i):
        if want_grad:
            h.eval_batch_device(d_par[i % nb].data_ptr(), B, d_nll.data_ptr(), d_tidx=d_tidx.data_ptr(), K=Kt,
                                d_grad=d_grad.data_ptr(), stream=stream.cuda_stream)
        else:
            h.eval_batch_device(d_par[i % nb].data_ptr(), B, d_nll.data_ptr(), stream=stream.cuda_stream)

    for i in range(warmup):
        step_device(i)
    barrier()
    clocks = ClockSampler(local) if main_run else None
    if clocks:
        clocks.start()
        time.sleep(0.25)
    launches0 = h.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    barrier()
    tc0 = time.perf_counter()
    evs[0].record(stream)
    for i in range(steps):
        step_device(i)
        evs[i + 1].record(stream)
    barrier()
    tc1 = time.perf_counter()
    launches = h.launch_count() - launches0
    total_ms = evs[0].elapsed_time(evs[-1])
    kern_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    clk = clocks.stop(tc0, tc1) if clocks else None
    assert torch.isfinite(d_nll).all().item(), "non-finite nll in the timed batch"
    last_nll = d_nll.cpu().numpy().copy()
    last_grad = d_grad[:2].cpu().numpy().copy() if want_grad else None

    # ---- end to end: host buffers in, host buffers out, through the public API
    def step_e2e(i):
        full = host[i % nb].numpy()                       # pinned; the H2D copy happens inside eval_batch
        nll, _, grad = h.eval_batch(full, tangent_idx=tix)
        if world > 1:                                     # every rank ends up with the whole job's nll vector
            t = torch.from_numpy(nll).to(dev)
            out = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            nll = torch.cat(out).cpu().numpy()
        return nll

    for i in range(min(warmup, 2)):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        step_e2e(i)
    barrier()
    e2e_s = time.perf_counter() - t0

    # ---- max over ranks
    tt = torch.tensor([total_ms, e2e_s, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = float(tt[0]), float(tt[1])

    value = world * B * steps / (total_ms * 1e-3)
    e2e = world * B * steps / e2e_s
    kms = float(np.mean(kern_ms))
    flop = ALG_FLOP[config]
    peak_tflops = env["peak_tflops"]
    achieved = flop * B / (kms * 1e-3) / 1e12
    info = h.tile_info()
    on_tile = a.kernel == 4 or (a.kernel == 0 and info["tile_ok"])
    kname = ("g3t::eval_tile_kernel" if on_tile else "g3::eval_thread_kernel") + ("<Dual<4>> (the dual pass carries the value)" if want_grad else "<double>")
    tr, cap = traffic_of(config)
    out = {
        "value": value, "unit": UNIT, "steps": steps, "warmup": warmup, "ms_per_step": total_ms / steps,
        "config": workload_cfg(config, B, len(base), len(opt), model.n_steps, world),
        "l2": "inputs larger than L2: %d distinct parameter batches (%.0f MB) rotate" % (nb, nb * B * P * 8 / 1e6),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(B * P * 8), "d2h_bytes_per_step": int(B * 8 * (1 + Kt))},
        "gpu_launches": int(launches),
        "roofline": {"bound": "fp64", "kernel": kname, "achieved": achieved, "peak": peak_tflops,
                     "unit": "TFLOP/s", "frac": achieved / peak_tflops if peak_tflops > 0 else None,
                     "peak_source": "DFMA-chain microbenchmark run live on this GPU (g3cuda_fp64_peak_tflops); "
                                    "MEASURED_PEAKS.json has no fp64 entry",
                     "alg_flop_per_eval": flop, "evals_per_launch": B, "launch_ms": kms,
                     "hbm": {"achieved": B * (P * 8 + 8) / (kms * 1e-3) / 1e9, "peak": env["peaks"].get("hbm_gbs"), "unit": "GB/s",
                             "alg_bytes_per_eval": P * 8 + 8},
                     "traffic": tr * B if (tr is not None and on_tile) else None,
                     "traffic_unit": ("bytes per launch: DRAM read + write per evaluation in the ncu capture %s x evaluations per launch" % cap)
                     if (tr is not None and on_tile) else None,
                     "tile": {"warps": info["warps"] + 1, "state_in_shared_memory": bool(info["state_in_smem"]) and not want_grad}},
    }
    if clk is not None:
        out["clocks"] = clk
    # ---- the timed path against the checker, on a slice of the last timed batch (never part of a timed region)
    if rank == 0:
        orc = load_oracle(model, params)
        cores = os.cpu_count() or 1
        big = sum(L * A_ * R for L, A_, R in model.stock_shapes) > 400
        nchk = 64 if big else 256
        chk = host[(steps - 1) % nb].numpy()[:nchk]
        g_nll, g_comp, _ = h.eval_batch(chk, want_comp=True)
        o_nll, o_comp, _ = orc.eval_batch(chk, want_comp=True, nthreads=cores)
        from oracle.parity import spot_check
        sc = spot_check(model, params, g_nll, g_comp, o_nll, o_comp)
        # the values the timed launches wrote are the ones that were checked (gradient runs: to rounding, another kernel)
        sc["timed_launch_vs_checked_call"] = float(np.max(np.abs(last_nll[:nchk] - g_nll) / np.abs(g_nll)))
        sc["within_contract"] = sc["within_contract"] and spot_check(model, params, last_nll[:nchk], g_comp, o_nll, o_comp)["within_contract"]
        if want_grad:
            _, _, og = orc.eval_batch(chk[:2], tangent_idx=tix, nthreads=cores)
            sc["max_rel_err_grad"] = float((np.abs(last_grad - og) / np.abs(og).max(axis=1, keepdims=True)).max())
            sc["within_contract"] = sc["within_contract"] and sc["max_rel_err_grad"] < 1e-8
        out["parity_spot_check"] = sc
        assert sc["within_contract"], f"{config}: results outside the parity contract: {sc}"
        if main_run and world == 1 and not a.no_cpu_baseline:
            out["cpu_baseline"] = cpu_reference(orc, model, params, config, batch_fn, tix=tix)
    return out


def sweep_c5(env, n_axis=1000):
    """BASELINE.json config 5 as specified: a likelihood-profile grid over two parameters, the rows sharded over the
    ranks by gadget3_b200.shard.eval_sharded (strong scaling: the grid is the whole job), the all-gather of the results
    over NCCL inside the timed region; host buffers in and out."""
    import torch
    from gadget3_b200.configs import CONFIGS, profile_sweep
    from gadget3_b200.run_cuda import g3_cuda_adfun
    from gadget3_b200.shard import eval_sharded, shard_range
    rank, world, local, dev, dist = env["rank"], env["world"], env["local"], env["dev"], env["dist"]
    model, params = CONFIGS["C5"]()
    fn = g3_cuda_adfun(model, params, device=local)
    Popt = profile_sweep(params, ["prey.K", "prey.pred.l50"], n_axis)
    full = fn.full_par(Popt)
    B = full.shape[0]
    lo, hi = shard_range(B, rank, world)
    fn.handle.eval_batch(full[lo:lo + 4096])               # warm-up: tables, workspaces
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    nll, _, _ = eval_sharded(lambda blk: fn.handle.eval_batch(blk), full, dist=dist if world > 1 else None, device=dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    assert nll.shape == (B,) and np.isfinite(nll).all()
    return {"metric": METRIC, "unit": UNIT,
            "config": {"workload": f"C5 profile sweep: {n_axis} x {n_axis} grid over prey.K x prey.pred.l50, {B} vectors in total, "
                                   f"rows sharded over {world} GPU(s) by shard.eval_sharded", "parallelism": f"batch-sharded x{world}"},
            "scaling": "strong", "e2e": {"value": B / float(dt[0]), "unit": UNIT, "seconds": float(dt[0]),
                                        "h2d_bytes_per_step": int((hi - lo) * full.shape[1] * 8), "d2h_bytes_per_step": int((hi - lo) * 8),
                                        "collective": "all_gather of nll over NCCL, inside the timed region" if world > 1 else "none"},
            "grid_min_nll": float(nll.min())}


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.batch <= 0:
        a.batch = DEFAULT_BATCH[a.config]

    if a.impl == "reference":
        if rank != 0:
            return 0
        from gadget3_b200.configs import CONFIGS, parameter_batch
        want_grad = a.config == "C3"
        model, params = CONFIGS[MODEL_OF[a.config]]()
        opt = params.optimised_indices()
        base = params.values()

        def batch_fn(n, seed):
            full = np.tile(base, (n, 1))
            full[:, opt] = parameter_batch(params, n, seed=seed)
            return full

        cores = os.cpu_count() or 1
        orc = load_oracle(model, params)
        tix = opt.astype(np.int32) if want_grad else None
        probe = batch_fn(2 * cores, 99)
        t = time.perf_counter(); orc.eval_batch(probe, tangent_idx=tix, nthreads=cores); per = (time.perf_counter() - t) / len(probe)
        n = int(min(max(4.0 / max(per, 1e-9), 2 * cores), a.batch))     # ~4 s per step
        n = max(cores, n // cores * cores)
        sample = batch_fn(n, 7)
        for _ in range(a.warmup):
            orc.eval_batch(sample[: max(cores, n // 8)], tangent_idx=tix, nthreads=cores)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            orc.eval_batch(sample, tangent_idx=tix, nthreads=cores)
        dt = time.perf_counter() - t0
        v = n * a.steps / dt
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_cfg(a.config, a.batch, len(base), len(opt), model.n_steps, max(world, a.gpus)),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n} vectors per step ({a.config}: a bounded sample of the workload's batch), {cores} threads"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return 0

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from gadget3_b200 import _lib
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    env = {"rank": rank, "world": world, "local": local, "dev": dev, "dist": dist, "peaks": peaks,
           "peak_tflops": _lib.load().g3cuda_fp64_peak_tflops(local)}

    m = measure(a.config, a.batch, a.steps, a.warmup, env, a, main_run=True)
    out = {"metric": METRIC, "value": m.pop("value"), "unit": m.pop("unit"), "n_gpus": world, "steps": m.pop("steps"),
           "warmup": m.pop("warmup"), "ms_per_step": m.pop("ms_per_step"), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
    out.update(m)
    if not a.no_extra:
        extra = []
        for c in ("C1", "C3", "C4", "C5"):
            if c == a.config:
                continue
            e = measure(c, DEFAULT_BATCH[c], EXTRA_STEPS[c], 3, env, a, main_run=False)
            e["metric"] = METRIC
            extra.append(e)
        if world > 1:
            extra.append(sweep_c5(env))
        out["extra_configs"] = extra
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
