#!/usr/bin/env python
"""Per-step comparison sharded vs single GPU in the outlier regime (debugging aid; run under torchrun)."""
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
    N, seed = (1 << 23) * G, 77
    Q, R = np.array([[0.04]]), np.array([[0.5]])
    mid, params = StochVolSSM(0.95, 1.0).device_functor(Q, R)
    Y = np.random.default_rng(3).standard_normal(12) * 1.2
    Y[3], Y[7] = 9.0, -8.5
    for T in (7, 8, 9):
        sh = ShardedSir(N, 1, 1, dtype="f32", trace_len=T)
        sh.eng.init_gaussian(np.zeros(1), np.array([[0.6]]), seed)
        sh.eng.set_observations(Y[:T].reshape(-1, 1))
        for k in range(T):
            sh.step(mid, params, 2.0, k, seed, k, k)
        rows = sh.eng.trace[:T].cpu().numpy()
        local = sh.eng.get_particles()[:, 0]
        gathered = [None] * G
        dist.all_gather_object(gathered, local)
        if rank == 0:
            ref = SirEngine(N, 1, 1, dtype="f32", trace_len=T)
            ref.init_gaussian(np.zeros(1), np.array([[0.6]]), seed)
            ref.run(mid, params, Y[:T].reshape(-1, 1), 2.0, seed, fused=False)
            rrows = ref.trace[:T].cpu().numpy()
            xref = ref.get_particles()[:, 0]
            xsh = np.concatenate(gathered)
            bad = np.flatnonzero(xsh != xref)
            print(f"T={T}: identical {np.mean(xsh == xref) * 100:.4f}%  first bad {bad[:5]}  n_bad {bad.size}")
            for k in range(T):
                print("   step", k, "sharded", rows[k, :6], "\n          single ", rrows[k, :6])
            if bad.size:
                i = bad[0]
                print("   around first bad:", xsh[i - 2:i + 3], xref[i - 2:i + 3], " unique values sharded/single:",
                      np.unique(xsh).size, np.unique(xref).size)
            del ref
            torch.cuda.empty_cache()
        sh.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
