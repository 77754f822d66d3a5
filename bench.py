#!/usr/bin/env python
"""bench.py — NLP evals/sec (constraints + sparse Jacobian, fp64), Anymal batch.

A "step" is one batched evaluation (g and all CSR Jacobian values) of B
independent Anymal fly-trot / Block-terrain problem instances
(BASELINE.json configs[1], B = 4096 per GPU, weak scaling: every rank owns its
own 4096 instances, no data-path collective).

  python bench.py [--gpus N] [--steps K] [--warmup W]          # CUDA arm
  python bench.py --impl reference [...]                        # CPU restatement arm (oracle, all host cores)
  torchrun --nproc-per-node N bench.py --gpus N ...             # N > 1

Prints ONE JSON line (rank 0).  See DESIGN.md §6 for how every field is measured.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = "anymal_trot_block"   # --workload overrides (other BASELINE configs; not the driver's line)
BATCH_PER_GPU = 4096
DESCR = {
    "anymal_trot_block": "BASELINE configs[1]: Anymal fly-trot C1, Block terrain, T=2.0 s",
    "biped_walk_stairs": "BASELINE configs[2]: Biped walk C0, Stairs, fpowr recipe, T=2.0 s",
    "hyq_gallop_gap": "BASELINE configs[3]: HyQ gallop C4, Gap terrain, phase durations optimised, T=2.0 s",
    "anymal_trot_mixed": "BASELINE configs[4], one GPU's shard: Anymal fly-trot C1, terrains drawn from Slope/Chimney/Gap per instance",
    "hopper": "BASELINE configs[0]: Monoped hopper, FlatGround, T=2.0 s",
}
METRIC = "nlp_evals_per_sec"
UNIT = "evals/s"


def read_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([s.strip() for s in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx = float(s[1])
                for nm, v in zip(names, s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_problem():
    import towr_b200 as tb
    spec = tb.make_formulation(WORKLOAD).to_spec()
    return tb, spec, tb.Problem(spec)


def kernel_breakdown(p, B):
    """Per-kernel durations from a second, serialised pass (TWB_PROFILE=1: every kernel on one stream, CUDA events
    around each launch) in a child process, with each kernel's own algorithmic bytes.  The timed region of the main
    pass runs the three output kernels concurrently on three streams, so their durations are not separable there."""
    env = dict(os.environ, TWB_PROFILE="1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--quick", "--steps", "30", "--warmup", "5",
                              "--workload", WORKLOAD, "--batch", str(B)],
                             capture_output=True, text=True, timeout=300, env=env).stderr
    except Exception:
        return None
    rows = {}
    for nm, r0, nr in p.constraint_sets():
        rows[nm] = (r0, nr)
    rp = p.row_ptr()
    def share(pred):
        nz = sum(int(rp[r0 + nr] - rp[r0]) for nm, (r0, nr) in rows.items() if pred(nm))
        mm = sum(nr for nm, (r0, nr) in rows.items() if pred(nm))
        return 8 * B * (nz + mm)
    alg = {"DynOut": share(lambda n: n == "dynamic"), "RomOut": share(lambda n: n.startswith("rangeofmotion")),
           "NodeOut": share(lambda n: n != "dynamic" and not n.startswith(("rangeofmotion", "totalduration"))),
           "RomNodeOut": share(lambda n: n != "dynamic" and not n.startswith("totalduration")),
           "TransposeIn": 2 * 8 * B * p.n, "TransposeOut": 2 * 8 * B * p.m}
    res = {}
    for line in out.splitlines():
        parts = line.split()
        if line.startswith("[twb profile]") and "avg" in parts:
            name, us = parts[2], float(parts[parts.index("avg") + 1])
            res[name] = {"avg_us": us}
            if name == "TransposeIn":
                res[name]["note"] = "first kernel of a step: includes the host launch gap of the serialised pass (ncu: 13 us)"
                continue
            if name in alg and us > 0:
                res[name]["algorithmic_bytes"] = alg[name]
                res[name]["gbs"] = alg[name] / (us * 1e-6) / 1e9
    return res or None


def host_threads():
    """Host threads the CPU legs use: every core this process may run on (torchrun exports OMP_NUM_THREADS=1,
    which must not throttle the CPU arm — the thread count is passed to the oracle explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_arm(spec, problem, X, threads, target_seconds=12.0):
    """Times the CPU restatement (oracle) on a bounded sample of the same iterates: whole passes over the batch
    (at most 4096 instances each) until ~target_seconds of CPU work have been done."""
    import oracle_lib
    sample = min(len(X), 4096)
    oracle_lib.batch_eval(spec, X[:min(sample, 4 * threads)], threads=threads)      # warm-up (thread pool, page faults)
    done, t0 = 0, time.perf_counter()
    while True:
        r = oracle_lib.batch_eval(spec, X[:sample], threads=threads)
        assert r["rc"] == 0
        done += sample
        dt = time.perf_counter() - t0
        if dt >= target_seconds or done >= 64 * sample:
            return done / dt, done, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle_lib
    from towr_b200.configs import synthetic_iterates_fast
    tb, spec, p = make_problem()
    threads = host_threads()
    X = synthetic_iterates_fast(p, 2048)
    # each step: a bounded sample of the workload sized for ~1.5 s of CPU time
    t0 = time.perf_counter(); oracle_lib.batch_eval(spec, X[:max(threads, 8)], threads=threads)
    per_eval = (time.perf_counter() - t0) / max(threads, 8)
    sample = int(max(threads, min(len(X), 1.5 / per_eval)))
    for _ in range(args.warmup):
        oracle_lib.batch_eval(spec, X[:sample], threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_lib.batch_eval(spec, X[:sample], threads=threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD} ({DESCR.get(WORKLOAD, WORKLOAD)}), n={p.n} m={p.m} nnz={p.nnz}",
                   "note": "reference arm = CPU restatement of towr's evaluation (oracle port; towr itself needs Eigen+ifopt, absent here)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} instances per step, OpenMP over instances"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local):
    """Multi-GPU runs: pin this rank's host threads to the CPUs next to its GPU (NVML's CPU affinity mask), so that the
    pinned host buffers of the end-to-end leg are first-touched on the GPU's NUMA node and the D2H copies of the ranks do
    not all cross one socket link.  Returns the number of CPUs bound to (0: left alone)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def run_cuda(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from towr_b200.configs import synthetic_iterates_fast

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (towr_b200 has no CPU evaluation path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = bind_to_gpu_numa_node(local) if world > 1 and not os.environ.get("TWB_NO_NUMA_BIND") else 0
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tb, spec, p = make_problem()
    B = args.batch or BATCH_PER_GPU
    batch = p.batch(B, device=local)
    if WORKLOAD == "anymal_trot_mixed":
        batch.set_terrains(np.random.default_rng(7 + rank).choice([tb.SLOPE, tb.CHIMNEY, tb.GAP], B).astype(np.int32))

    # ---- inputs: a ring of distinct iterate sets, together larger than L2 (126 MB)
    n_ring = 8
    Xh = synthetic_iterates_fast(p, B, seed=1234 + rank)
    ring = []
    for r in range(n_ring):
        Xr = Xh if r == 0 else synthetic_iterates_fast(p, B, seed=1234 + rank + 1000 * r)
        ring.append(torch.from_numpy(Xr).to(dev))
    g = torch.empty((B, p.m), dtype=torch.float64, device=dev)
    jac = torch.empty((B, p.nnz), dtype=torch.float64, device=dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev)
    flags = tb.EVAL_G | tb.EVAL_JAC

    def step(i):
        batch.eval_device(ring[i % n_ring], g=g, jac=jac, status=status, flags=flags, stream=stream)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    sampler = ClockSampler(local) if rank == 0 and not args.quick else None
    if sampler:
        sampler.start()
    for i in range(max(args.warmup, 3)):
        step(i)
    sync_all()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    sync_all()
    ev[0].record(stream)
    for i in range(args.steps):
        step(i)
        ev[i + 1].record(stream)
    sync_all()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_step = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps))
    if sampler:
        # the timed region lasts a few milliseconds; keep the same loop running ~1 s more so that nvidia-smi (one
        # sample per ~60 ms) sees the clocks under this load; these steps are not part of any reported number
        t_end = time.perf_counter() + 1.0
        i = 0
        while time.perf_counter() < t_end:
            step(i); i += 1
            if i % 64 == 0:
                torch.cuda.synchronize(dev)
        torch.cuda.synchronize(dev)
        sampler.stop_flag = True
    assert int(status.sum().item()) == 0, "non-finite values flagged"
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * B * args.steps / (total_ms * 1e-3)

    if args.quick:
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": total_ms / args.steps, "ms_best": per_step[0],
                              "frac": 8 * (p.n + p.m + p.nnz) * B / (total_ms / args.steps * 1e-3) / 1e9 / read_peak()[0]}), flush=True)
        return
    # ---- end to end through the host-pointer C ABI call: pinned host buffers, H2D + D2H inside the timed region
    xs_pinned = torch.from_numpy(Xh).pin_memory()
    out = {"g": torch.empty((B, p.m), dtype=torch.float64).pin_memory().numpy(),
           "jac": torch.empty((B, p.nnz), dtype=torch.float64).pin_memory().numpy(),
           "status": torch.empty(B, dtype=torch.int32).pin_memory().numpy()}
    x_np = xs_pinned.numpy()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        batch.eval_host(x_np, flags=flags, out=out)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        batch.eval_host(x_np, flags=flags, out=out)      # synchronises internally; result is on the host
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(te.item())
    h2d = B * p.n * 8
    d2h = B * (p.m + p.nnz) * 8 + B * 4

    if rank == 0:
        import oracle_lib
        peak, peak_src = read_peak()
        bytes_per_eval = 8 * (p.n + p.m + p.nnz)
        avg_kernel_ms = total_ms / args.steps       # one evaluation (all its kernels) per step
        achieved = bytes_per_eval * B / (avg_kernel_ms * 1e-3) / 1e9
        threads = host_threads()
        cpu_value, cpu_sample, cpu_dt = cpu_arm(spec, p, Xh, threads)
        kernels = kernel_breakdown(p, B)
        dominant = None
        if kernels:
            cand = [(k, v) for k, v in kernels.items() if "gbs" in v and k.endswith("Out") and k != "TransposeOut"]
            if cand:
                k, v = max(cand, key=lambda kv: kv[1]["avg_us"])
                dominant = {"name": k, "avg_launch_ms": v["avg_us"] * 1e-3, "algorithmic_bytes_per_launch": v["algorithmic_bytes"],
                            "achieved": v["gbs"], "frac": v["gbs"] / peak, "how": "serialised pass, CUDA events around each launch"}
        traffic, traffic_kernels = None, {}
        tpath = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
        if os.path.exists(tpath) and WORKLOAD == "anymal_trot_block" and B == BATCH_PER_GPU:
            try:
                tj = json.load(open(tpath))
                traffic, traffic_kernels = tj.get("dram_bytes_per_launch"), tj.get("kernels", {})
            except Exception:
                traffic = None
        if dominant and dominant["name"] in traffic_kernels:
            dominant["traffic"] = traffic_kernels[dominant["name"]]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{WORKLOAD} ({DESCR.get(WORKLOAD, WORKLOAD)}), "
                                   f"{B} instances per GPU, n={p.n} m={p.m} nnz={p.nnz}",
                       "batch_per_gpu": B, "outputs": "g[B][m] + jac[B][nnz] (CSR values), fp64",
                       "l2": f"inputs rotate over {n_ring} distinct iterate sets ({n_ring * B * p.n * 8 / 1e6:.0f} MB) and each step "
                             f"writes {B * (p.m + p.nnz) * 8 / 1e6:.0f} MB of outputs (> 126 MB L2)",
                       "iterates": "x0 + sigma*N(0,1), sigma 0.05 pos / 0.2 vel / 10 N force (SURVEY §8d)"},
            "ms_per_step_median": per_step[len(per_step) // 2], "ms_per_step_best": per_step[0],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_per_eval * B, "avg_launch_ms": avg_kernel_ms,
                         "kernel": "whole evaluation: TransposeIn -> RomNodeOut | DynOut (two streams) -> TransposeOut; "
                                   "CUDA events on the launching stream around every step of the timed region",
                         "dominant_kernel": dominant},
            "kernels": kernels,
            "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{cpu_sample} evaluations of instances of the same batch, OpenMP over instances, {cpu_dt:.1f} s"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "twb_batch_eval_host (pinned host buffers)",
                    "host_cpus_bound_per_rank": numa_cpus},
            "gpu_launches": args.steps * batch.launches_per_eval(flags),
            "clocks": sampler.summary() if sampler else None,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--quick", action="store_true", help="kernel timing only (tuning sweeps): skip e2e and the CPU arm")
    ap.add_argument("--workload", default=None, help="recipe name of towr_b200.configs (default: BASELINE configs[1])")
    ap.add_argument("--batch", type=int, default=0, help="instances per GPU (default 4096)")
    args = ap.parse_args()
    global WORKLOAD
    if args.workload:
        WORKLOAD = args.workload
    import __graft_entry__ as ge
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        ge.build(quiet=True)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
