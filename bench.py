#!/usr/bin/env python
"""bench.py -- throughput of the reversal-distance hot path on the workload BASELINE.json quotes:
config 3 = all 499,500 pairs of 1000 synthetic genomes x 2000 signed blocks (seed 7).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one pass of the path over the whole pair list (at N > 1 each rank owns a contiguous
1/N slice of the upper triangle; no collective on the data path; strong scaling of a fixed job).

  value          pairs/s, kernel only, inputs resident in HBM, per-step CUDA events on the launch
                 stream, L2 flushed between steps, max over ranks
  e2e            pairs/s through the C ABI call a GRIMM binding makes (grsr_dist_matrix_shard) with
                 HOST buffers: H2D of the genome table, signed inverses, pair generation, kernel,
                 D2H of the distances all inside the timed region
  roofline       algorithmic bytes (8n+4 per pair) / kernel time vs the
                 measured HBM peak of MEASURED_PEAKS.json
  cpu_baseline   the UNMODIFIED reference (oracle/_ref/libgrimmref.so) on all host cores, on a
                 bounded sample of the same pair list (rank 0, N = 1 only)

--impl reference times that reference CPU implementation alone (same metric / config keys).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from grsr_b200 import gen  # noqa: E402

G_CFG, N_CFG, SEED_CFG = 1000, 2000, 7
METRIC = "reversal_distance_pairs_per_sec"
UNIT = "pairs/s"
BYTES_PER_PAIR = 8 * N_CFG + 4          # read pi and gamma as int32, write one int32 distance
I32P = ctypes.POINTER(ctypes.c_int)


def workload_config(extra=None):
    cfg = {
        "workload": "config3: all-pairs reversal distance, 1000 genomes x 2000 signed blocks "
                    "(499500 pairs), synthetic seed 7, circular-aligned (-C -m)",
        "genomes": G_CFG, "genes": N_CFG, "pairs": G_CFG * (G_CFG - 1) // 2,
        "sharding": "contiguous slices of the column-by-column upper triangle, one per rank, no collective",
    }
    if extra:
        cfg.update(extra)
    return cfg


def make_genomes():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    g = gen.random_genomes(G_CFG, N_CFG, SEED_CFG)
    # -C: align every genome on the first gene of genome 1 (host preprocessing of the CLI,
    # reference G/circ_align.c:50); genome 1 is the identity so this is a rotation / flip per row
    out = g.copy()
    a = int(g[0, 0])
    for r in range(1, G_CFG):
        row = g[r]
        pos = int(np.nonzero(np.abs(row) == abs(a))[0][0])
        if row[pos] == a:
            out[r] = np.concatenate([row[pos:], row[:pos]])
        else:
            out[r] = np.concatenate([-row[:pos + 1][::-1], -row[pos + 1:][::-1]])
    return np.ascontiguousarray(out, dtype=np.int32)


# ---- reference CPU arm ---------------------------------------------------------------------------

def load_cpu_reference():
    """(kind, callable(genomes, pairs, threads) -> distances).  Prefers the unmodified reference
    compiled in place; falls back to the oracle port (single thread per process)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "libgrimmref.so")
    if os.path.exists(ref):
        lib = ctypes.CDLL(ref)

        def run(genomes, pairs, threads):
            out = np.zeros(len(pairs), dtype=np.int32)
            lib.ref_dist_pairs(genomes.ctypes.data_as(I32P), genomes.shape[1], pairs.ctypes.data_as(I32P),
                               ctypes.c_long(len(pairs)), out.ctypes.data_as(I32P), 1, threads)
            return out
        return "reference", run
    port = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(port):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True)
    lib = ctypes.CDLL(port)

    def run1(genomes, pairs, threads):
        out = np.zeros(len(pairs), dtype=np.int32)
        ths = []
        for t in range(threads):
            lo, hi = len(pairs) * t // threads, len(pairs) * (t + 1) // threads
            sl = np.ascontiguousarray(pairs[lo:hi])
            o = out[lo:hi]
            ths.append(threading.Thread(target=lib.orc_dist_pairs, args=(
                genomes.ctypes.data_as(I32P), genomes.shape[1], sl.ctypes.data_as(I32P),
                ctypes.c_long(hi - lo), o.ctypes.data_as(I32P), 1)))
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        return out
    return "port", run1


def cpu_sample_rate(genomes, pairs, seconds: float):
    """Time the CPU reference on a bounded prefix of the pair list sized for about `seconds`."""
    kind, run = load_cpu_reference()
    cores = os.cpu_count() or 1
    probe = pairs[:min(len(pairs), 64 * cores)]
    t0 = time.perf_counter()
    run(genomes, probe, cores)
    dt = max(time.perf_counter() - t0, 1e-6)
    want = int(min(len(pairs), max(len(probe), len(probe) / dt * seconds)))
    sample = np.ascontiguousarray(pairs[:want])
    t0 = time.perf_counter()
    d = run(genomes, sample, cores)
    dt = time.perf_counter() - t0
    return {"value": len(sample) / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "first %d of %d pairs of the workload, %.1f s on %d host threads"
                      % (len(sample), len(pairs), dt, cores)}, d, sample


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    genomes = make_genomes()
    j, i = np.tril_indices(G_CFG, k=-1)
    pairs = np.stack([i, j], axis=1).astype(np.int32)   # column-by-column triangle, as the GPU arm
    kind, run = load_cpu_reference()
    cores = os.cpu_count() or 1
    # size one step for a few seconds so that steps+warmup stay within minutes
    probe = np.ascontiguousarray(pairs[:64 * cores])
    t0 = time.perf_counter()
    run(genomes, probe, cores)
    rate = len(probe) / max(time.perf_counter() - t0, 1e-6)
    budget = 120.0 / max(1, args.steps + args.warmup)
    per_step = int(min(len(pairs), max(len(probe), rate * min(budget, 10.0))))
    sample = np.ascontiguousarray(pairs[:per_step])
    for _ in range(args.warmup):
        run(genomes, sample, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run(genomes, sample, cores)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic",
        "config": workload_config({"reference_step": "first %d pairs of the list per step" % per_step}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d pairs per step x %d steps on %d host threads" % (per_step, args.steps, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- clocks --------------------------------------------------------------------------------------

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, ln in self.lines:
            if t0 is not None and not (t0 <= ts <= t1 + 0.2):
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- GPU arm -------------------------------------------------------------------------------------

def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from grsr_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the path has no CPU implementation")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    genomes = make_genomes()
    P = G_CFG * (G_CFG - 1) // 2
    lo, hi = P * rank // world, P * (rank + 1) // world
    cnt = hi - lo

    # resident inputs
    stream = torch.cuda.current_stream().cuda_stream
    d_gen = torch.from_numpy(genomes).to(dev)
    d_sinv = torch.empty_like(d_gen)
    d_pairs = torch.empty((cnt, 2), dtype=torch.int32, device=dev)
    d_out = torch.empty((cnt,), dtype=torch.int32, device=dev)
    capi.signed_inverse_dev(d_gen.data_ptr(), G_CFG, N_CFG, d_sinv.data_ptr(), stream)
    capi.triangle_pairs_dev(G_CFG, lo, cnt, d_pairs.data_ptr(), stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step_dev():
        capi.dist_pairs_dev(d_gen.data_ptr(), d_sinv.data_ptr(), N_CFG, d_pairs.data_ptr(), cnt,
                            d_out.data_ptr(), 1, stream)

    for _ in range(max(args.warmup, 3)):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_region0 = time.perf_counter()
    for a, b in ev:
        flush.fill_(1)                 # evict L2 between timed steps (outside the event pair)
        a.record()
        step_dev()
        b.record()
    barrier()
    t_region1 = time.perf_counter()
    ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = max_over_ranks(float(sum(ms)))
    kernel_ms = float(np.mean(ms))
    value = P * args.steps / (total_ms / 1e3)

    # end to end through the host-buffer C ABI
    h_gen = torch.from_numpy(genomes).pin_memory()
    h_out = torch.empty((cnt,), dtype=torch.int32).pin_memory()
    g_np, o_np = h_gen.numpy(), h_out.numpy()
    for _ in range(max(args.warmup, 3)):
        capi.dist_matrix_shard(g_np, rank, world, device=local, out=o_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        capi.dist_matrix_shard(g_np, rank, world, device=local, out=o_np)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    clocks = sampler.stop(t_region0, t_region1) if rank == 0 else None
    # the two paths must agree before any number is reported
    if not np.array_equal(o_np, d_out.cpu().numpy()):
        raise SystemExit("device-resident and host-buffer paths disagree")
    e2e_value = P * args.steps / e2e_s

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        achieved = cnt * BYTES_PER_PAIR / (kernel_ms / 1e3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "dist_kernel_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic,
                    "kernel": "grsr::dist_walk_shared_kernel (+ walk_order_kernel, dist_deferred_kernel on the deferred pairs)",
                    "kernel_ms": kernel_ms,
                    "algorithmic_bytes_per_pair": BYTES_PER_PAIR, "pairs_per_launch": cnt, "peak_source": peak_src}
        cpu = None
        if world == 1 and not args.no_cpu:
            pairs = capi.triangle_pairs(G_CFG)          # same flat order as the GPU result
            cpu, d_cpu, sample = cpu_sample_rate(genomes, pairs, seconds=12.0)
            if not np.array_equal(d_cpu, o_np[:len(sample)]):
                raise SystemExit("GPU distances differ from the CPU reference on the baseline sample")
            cpu["checked"] = "GPU distances equal the CPU reference on all %d sample pairs" % len(sample)
        scen = scenario_leg(capi) if (world == 1 and not args.no_cpu and not args.no_scenario) else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config({"l2": "flushed between timed steps with a 256 MiB write; per-step CUDA events",
                                       "threads_per_block": os.environ.get("GRSR_THREADS", "auto")}),
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(genomes.nbytes) * world,
                    "d2h_bytes_per_step": int(P * 4), "ms_per_step": e2e_s / args.steps * 1e3,
                    "api": "grsr_dist_matrix_shard (host buffers, pinned)"},
            "gpu_launches": 4 * args.steps * world, "clocks": clocks, "scenario": scen,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---- scenario leg (secondary metric of BASELINE.json: scenario reversals/s) ---------------------------

def _ref_scenario_worker(args):
    """One reference scenario in a worker process (the reference prints through a global FILE*)."""
    src, dst = args
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libgrimmref.so"))
    n = len(src)
    steps = np.zeros(2 * (n + 2), dtype=np.int32)
    k = lib.ref_scenario(src.ctypes.data_as(I32P), dst.ctypes.data_as(I32P), n, steps.ctypes.data_as(I32P), n + 2)
    return steps[:2 * max(k, 0)].reshape(-1, 2)


def scenario_leg(capi):
    """Reversals/s of grsr_scenario_batch (host buffers in, step lists out) on two bounded samples:
    config-4-sized random pairs, and hurdle/fortress-rich pairs (where the reference brute-forces
    every breakpoint pair); the latter is also timed on the reference CPU and compared step by step."""
    import random
    from concurrent.futures import ProcessPoolExecutor
    out = {}
    rng = random.Random(4)
    n, k = 5000, 148
    src = np.array([gen.random_signed_perm(n, rng) for _ in range(k)], dtype=np.int32)
    dst = np.tile(np.arange(1, n + 1, dtype=np.int32), (k, 1))
    capi.scenario_batch(src[:2], dst[:2])
    t0 = time.perf_counter()
    steps = capi.scenario_batch(src, dst)
    dt = time.perf_counter() - t0
    tot = sum(len(x) for x in steps)
    out["random_n5000"] = {"pairs": k, "genes": n, "reversals": tot, "seconds": dt, "value": tot / dt,
                           "unit": "reversals/s", "roofline_frac_hbm": tot / dt * (12 * n + 16) / 6554.6e9}
    n, k = 300, 2368
    src = np.array([gen.hurdle_rich_perm(n, 9000 + s) for s in range(k)], dtype=np.int32)
    dst = np.tile(np.arange(1, n + 1, dtype=np.int32), (k, 1))
    t0 = time.perf_counter()
    steps = capi.scenario_batch(src, dst)
    dt = time.perf_counter() - t0
    tot = sum(len(x) for x in steps)
    leg = {"pairs": k, "genes": n, "reversals": tot, "seconds": dt, "value": tot / dt, "unit": "reversals/s"}
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libgrimmref.so")):
        cores = os.cpu_count() or 1
        m = min(k, 2 * cores)
        t0 = time.perf_counter()
        with ProcessPoolExecutor(max_workers=cores) as ex:
            ref = list(ex.map(_ref_scenario_worker, [(src[q], dst[q]) for q in range(m)]))
        rdt = time.perf_counter() - t0
        if not all(np.array_equal(a, b) for a, b in zip(ref, steps[:m])):
            raise SystemExit("GPU scenarios differ from the CPU reference on the scenario sample")
        rtot = sum(len(x) for x in ref)
        leg["cpu_baseline"] = {"value": rtot / rdt, "unit": "reversals/s", "cores": cores, "kind": "reference",
                               "sample": "first %d pairs, one process per pair on %d cores, %.1f s; step lists equal"
                                         % (m, cores, rdt)}
    out["hurdle_rich_n300"] = leg
    # one hurdle-rich pair alone (the command-line case): the library scores each step's candidates
    # across the whole GPU (grsr_stepwise.cuh) instead of one after another inside one block
    n = 300
    one_src = np.array(gen.hurdle_rich_perm(n, 1), dtype=np.int32)
    one_dst = np.arange(1, n + 1, dtype=np.int32)
    capi.scenario_batch(one_src, one_dst)
    t0 = time.perf_counter()
    one = capi.scenario_batch(one_src, one_dst)[0]
    dt = time.perf_counter() - t0
    leg = {"pairs": 1, "genes": n, "reversals": len(one), "seconds": dt, "value": len(one) / dt,
           "unit": "reversals/s"}
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libgrimmref.so")):
        t0 = time.perf_counter()
        ref1 = _ref_scenario_worker((one_src, one_dst))
        rdt = time.perf_counter() - t0
        if not np.array_equal(ref1, one):
            raise SystemExit("GPU scenario differs from the CPU reference on the single-pair sample")
        leg["cpu_baseline"] = {"value": len(ref1) / rdt, "unit": "reversals/s", "cores": 1, "kind": "reference",
                               "sample": "the same pair, one process, %.2f s; step lists equal" % rdt}
    out["single_hurdle_rich_pair_n300"] = leg
    # config-4 size: the unit of work of a hurdle-rich scenario is one full distance evaluation of
    # a trial permutation (the reference makes ~1e7 of them per pair at n = 5000); score 16384
    # candidate reversals of one hurdle-rich pair in one call and compare a sample with the reference
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    n = 5000
    src5 = np.array(gen.hurdle_rich_perm(n, 4), dtype=np.int32)
    dst5 = np.arange(1, n + 1, dtype=np.int32)
    rng = np.random.default_rng(4)
    ci = rng.integers(0, n, size=16384)
    cj = rng.integers(0, n, size=16384)
    cands = np.stack([np.minimum(ci, cj), np.maximum(ci, cj)], axis=1).astype(np.int32)
    capi.trial_distances(src5, dst5, cands[:64])
    t0 = time.perf_counter()
    d2 = capi.trial_distances(src5, dst5, cands)
    dt = time.perf_counter() - t0
    leg = {"candidates": len(cands), "genes": n, "seconds": dt, "value": len(cands) / dt,
           "unit": "evaluations/s", "roofline_frac_hbm": len(cands) / dt * 8 * n / 6554.6e9}
    ref = os.path.join(ROOT, "oracle", "_ref", "libgrimmref.so")
    if os.path.exists(ref):
        lib = ctypes.CDLL(ref)
        cores = os.cpu_count() or 1
        m = 64 * cores
        trials = np.tile(src5, (m, 1))
        for q in range(m):
            i, j = cands[q]
            trials[q, i:j + 1] = -src5[i:j + 1][::-1]
        table = np.ascontiguousarray(np.vstack([trials, dst5[None, :]]), dtype=np.int32)
        prs = np.stack([np.arange(m), np.full(m, m)], axis=1).astype(np.int32)   # gamma = trial, pi = dest
        outr = np.zeros(m, dtype=np.int32)
        t0 = time.perf_counter()
        lib.ref_dist_pairs(table.ctypes.data_as(I32P), n, prs.ctypes.data_as(I32P), ctypes.c_long(m),
                           outr.ctypes.data_as(I32P), 1, cores)
        rdt = time.perf_counter() - t0
        if not np.array_equal(outr, d2[:m]):
            raise SystemExit("GPU trial distances differ from the CPU reference")
        leg["cpu_baseline"] = {"value": m / rdt, "unit": "evaluations/s", "cores": cores, "kind": "reference",
                               "sample": "first %d candidates on %d host threads, %.3f s; distances equal" % (m, cores, rdt)}
    out["trial_evaluations_n5000"] = leg
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="grsr", choices=["grsr", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline and scenario legs")
    ap.add_argument("--no-scenario", action="store_true", help="skip the scenario leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
