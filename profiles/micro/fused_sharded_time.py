#!/usr/bin/env python
"""Per-step time of the fused sharded step (one launch, one exchange) against the single-GPU fused step at the same
shard size.  Run under torchrun with any number of ranks (1 = the kernel's own overhead without peers):
    python -m torch.distributed.run --nproc-per-node 1 --master-addr 127.0.0.1 profiles/micro/fused_sharded_time.py 25 24 23"""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)


def main():
    import torch
    from particle_filters_b200 import StochVolSSM
    from particle_filters_b200.engine import SirEngine
    from particle_filters_b200.sharded import ShardedSir, init_distributed
    dist = init_distributed()
    rank, G = dist.get_rank(), dist.get_world_size()
    Q, R = np.array([[0.04]]), np.array([[1.0]])
    mid, params = StochVolSSM(0.95, 1.0).device_functor(Q, R)
    Y = np.random.default_rng(3).standard_normal(64) * 1.2
    K, seed, thresh = 400, 42, 2.0
    for lg in [int(a) for a in sys.argv[1:]] or [25]:
        n = 1 << lg                      # per shard
        sh = ShardedSir(n * G, 1, 1, dtype="f32", trace_len=8)
        sh.eng.init_gaussian(np.zeros(1), np.array([[0.6]]), seed)
        sh.eng.set_observations(Y.reshape(-1, 1))
        res = {}
        for mode in ("fused", "three-launch"):
            if mode == "fused":
                sh.fused_begin(mid, params, thresh, 0, seed, 0)
            step = (lambda k: sh.fused_step(mid, params, thresh, k % 64, seed, k, 4)) if mode == "fused" else \
                   (lambda k: sh.step(mid, params, thresh, k % 64, seed, k, 4))
            for k in range(1, 101):
                step(k)
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(101, 101 + K):
                step(k)
            e1.record(); torch.cuda.synchronize()
            if mode == "fused":
                sh.fused_end(thresh, seed, 100 + K, 4)
            ms = torch.tensor([e0.elapsed_time(e1) / K], dtype=torch.float64, device="cuda")
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            res[mode] = float(ms.item())
        sh.check(); sh.close()
        single = None
        if rank == 0:
            eng = SirEngine(n, 1, 1, dtype="f32", trace_len=64)
            eng.init_gaussian(np.zeros(1), np.array([[0.6]]), seed)
            eng.set_observations(Y.reshape(-1, 1))
            eng.enqueue(mid, params, 64, thresh, seed); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                eng.enqueue(mid, params, 64, thresh, seed)
            e1.record(); torch.cuda.synchronize()
            single = e0.elapsed_time(e1) / 256
            del eng; torch.cuda.empty_cache()
            print(f"[fused_sharded_time] G={G} shard 2^{lg}: fused sharded {res['fused']*1e3:.1f} us, three-launch sharded "
                  f"{res['three-launch']*1e3:.1f} us, single-GPU fused at the shard size {single*1e3:.1f} us", flush=True)
        dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
