"""BASELINE configs[4] as a solver loop at full size: 65 536 Anymal multi-start instances (mixed Slope / Chimney / Gap terrains,
goal-randomised) split over the ranks with shard_range; every rank walks its shard through `--iters` Levenberg-Marquardt
feasibility iterations on its own GPU (twb_batch_goal_instances_device once, then twb_batch_eval_device +
twb_batch_lm_step_device per iteration; nothing leaves the device), then the per-instance violation histories' first / last
entries and the status words are all-gathered over NCCL.  Timing: CUDA events around the loop, max over ranks.

    python scripts/solver_loop_multi.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
        scripts/solver_loop_multi.py --out gpurun_out/solver_loop_n8.json
"""
import argparse, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    ap = argparse.ArgumentParser()
    ap.add_argument("--total", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    import towr_b200 as tb
    from towr_b200.sharding import shard_range, shard_sizes
    from towr_b200.solver import BatchedLevenbergMarquardt
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = shard_range(args.total, rank, world)
    nb = hi - lo
    spec = tb.make_formulation("anymal_trot_mixed").to_spec()
    p = tb.Problem(spec)
    rng = np.random.default_rng(77)                       # the same global draw on every rank, sliced by the shard
    terr = rng.choice([tb.SLOPE, tb.CHIMNEY, tb.GAP], args.total).astype(np.int32)
    goals = np.column_stack([rng.uniform(1.0, 2.0, args.total), rng.uniform(-0.2, 0.2, args.total), np.full(args.total, 0.5),
                             np.zeros(args.total), np.zeros(args.total), rng.uniform(-0.2, 0.2, args.total)])
    noise = np.random.default_rng(1000 + rank).standard_normal((nb, p.n))
    bt = p.batch(nb, device=local)
    bt.set_terrains(terr[lo:hi])
    x0, xl, xu = bt.goal_instances_device(torch.from_numpy(goals[lo:hi]).to(dev))
    X = torch.minimum(torch.maximum(x0 + 0.01 * torch.from_numpy(noise).to(dev), xl), xu)
    lm = BatchedLevenbergMarquardt(bt, x_lower=xl, x_upper=xu)
    lm.run(X.clone(), 1)                                  # warm-up on a copy (graph capture, pattern upload)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier(); torch.cuda.synchronize(dev)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    a.record()
    hist = lm.run(X, args.iters)
    b.record()
    barrier()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # gather first / last violation and status of every instance (NCCL all_gather at N > 1)
    sizes = shard_sizes(args.total, world); pad = max(sizes)
    loc = torch.zeros((pad, 3), dtype=torch.float64, device=dev)
    loc[:nb, 0] = hist[0]; loc[:nb, 1] = hist[-1]; loc[:nb, 2] = lm.status.to(torch.float64)
    if world > 1:
        parts = [torch.empty_like(loc) for _ in range(world)]
        dist.all_gather(parts, loc)
        allv = torch.cat([parts[r][:sizes[r]] for r in range(world)]).cpu().numpy()
    else:
        allv = loc[:nb].cpu().numpy()
    if rank == 0:
        rec = {"workload": "anymal_trot_mixed: BASELINE configs[4], multi-start feasibility loop", "instances_total": args.total, "n_gpus": world,
               "instances_per_gpu": nb, "iterations": args.iters, "cg_iters_per_iteration": lm.cg_iters, "ms_total_max_over_ranks": ms,
               "lm_iterations_per_s": args.iters / (ms * 1e-3), "instance_iterations_per_s": args.total * args.iters / (ms * 1e-3),
               "gathered_instances": int(allv.shape[0]), "flagged": int((allv[:, 2] != 0).sum()),
               "violation_median_first_last": [float(np.median(allv[:, 0])), float(np.median(allv[:, 1]))],
               "fraction_improved": float((allv[:, 1] < allv[:, 0]).mean()),
               "collective": "all_gather of (first violation, last violation, status) per instance: " + ("nccl" if world > 1 else "none (one rank)")}
        print(json.dumps(rec), flush=True)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            json.dump(rec, open(args.out, "w"), indent=1)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
