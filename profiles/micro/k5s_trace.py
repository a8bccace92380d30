#!/usr/bin/env python
"""Where a step of the fused sharded kernel spends its time: per-CTA globaltimer stamps of the last two launches
(tuning build:  profiles/micro/build_variant.sh trace -DSMC_K5_TRACE;  SMC_LIB_PATH=profiles/micro/variants/libsmc_trace.so).
    RANK=0 WORLD_SIZE=1 LOCAL_RANK=0 MASTER_ADDR=127.0.0.1 MASTER_PORT=29541 python profiles/micro/k5s_trace.py 23"""
import ctypes as C, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)


def main():
    import torch
    from particle_filters_b200 import StochVolSSM, _lib
    from particle_filters_b200.sharded import ShardedSir, init_distributed
    dist = init_distributed()
    rank, G = dist.get_rank(), dist.get_world_size()
    mid, params = StochVolSSM(0.95, 1.0).device_functor(np.array([[0.04]]), np.array([[1.0]]))
    Y = np.random.default_rng(3).standard_normal(64) * 1.2
    seed, thresh = 42, 2.0
    lib = _lib.load()
    lib.smc_k5_trace_read.argtypes = [C.c_void_p]
    for lg in [int(a) for a in sys.argv[1:]] or [23]:
        n = 1 << lg
        sh = ShardedSir(n * G, 1, 1, dtype="f32", trace_len=8)
        sh.eng.init_gaussian(np.zeros(1), np.array([[0.6]]), seed)
        sh.eng.set_observations(Y.reshape(-1, 1))
        sh.fused_begin(mid, params, thresh, 0, seed, 0)
        for k in range(1, 301):
            sh.fused_step(mid, params, thresh, k % 64, seed, k, 4)
        torch.cuda.synchronize()
        buf = np.zeros(2 * 1024 * 16, dtype=np.uint64)
        lib.smc_k5_trace_read(buf.ctypes.data_as(C.c_void_p))
        sh.fused_end(thresh, seed, 300, 4)
        sh.check(); sh.close()
        tr = buf.reshape(2, 1024, 16).astype(np.int64)
        last = sh.stamp - 1                   # stamp of the last fused launch (fused_end took one more)
        cur, prv = tr[last & 1], tr[(last - 1) & 1]
        ncta = int(np.count_nonzero(cur[:, 0]))
        cur, prv = cur[:ncta], prv[:ncta]
        t0 = cur[:, 0].min()
        done_prev = prv[:, 6].max()
        done_cur = cur[:, 6].max()
        rel = lambda a: (a - t0) / 1e3
        if rank == 0:
            print(f"[k5s_trace] G={G} shard 2^{lg}: {ncta} CTAs; step period {(done_cur - done_prev) / 1e3:.1f} us")
            print(f"  previous launch's flag -> first CTA entry {(t0 - done_prev) / 1e3:6.1f} us;  entries spread "
                  f"{rel(cur[:, 0]).max():.1f} us")
            for name, i, j in (("flag wait", 0, 1), ("finalize + finish", 1, 2), ("CDF prefix", 2, 3), ("tiles", 3, 4),
                               ("CTA tail (records, fence)", 4, 5)):
                d = (cur[:, j] - cur[:, i]) / 1e3
                print(f"  {name:28s} median {np.median(d):6.1f}  min {d.min():6.1f}  max {d.max():6.1f} us")
            print(f"  tile loop starts  at {rel(cur[:, 3]).min():6.1f} .. {rel(cur[:, 3]).max():6.1f} us after the first entry")
            print(f"  tile loop ends    at {rel(cur[:, 4]).min():6.1f} .. {rel(cur[:, 4]).max():6.1f} us")
            print(f"  last ticket at {rel(cur[:, 5]).max():6.1f} us, flag raised at {rel(np.array([done_cur]))[0]:6.1f} us")
            if os.environ.get("K5S_TRACE_DUMP"):
                np.save(os.path.join(REPO, "gpurun_out", f"k5s_trace_{lg}.npy"), cur)
            lc = int(np.argmax(cur[:, 6]))
            tl = cur[lc]
            print(f"  last CTA (block {lc}): ticket {rel(tl[5:6])[0]:.1f}; loads issued +{(tl[10]-tl[5])/1e3:.1f}; combine +{(tl[11]-tl[10])/1e3:.1f}; "
                  f"moments + publish +{(tl[12]-tl[11])/1e3:.1f}; zero + ctrl +{(tl[13]-tl[12])/1e3:.1f}; fence +{(tl[14]-tl[13])/1e3:.1f}; "
                  f"flag +{(tl[6]-tl[14])/1e3:.1f} us")
            tiles = (cur[:, 4] - cur[:, 3]) / 1e3
            smid = cur[:, 7]
            per_sm = {}
            for b in range(ncta):
                per_sm.setdefault(int(smid[b]), []).append(float(tiles[b]))
            sm_mean = np.array([np.mean(v) for v in per_sm.values()])
            print(f"  SMs used {len(per_sm)}, CTAs per SM {min(len(v) for v in per_sm.values())}..{max(len(v) for v in per_sm.values())}; "
                  f"per-SM mean tile-loop time {sm_mean.min():.0f}..{sm_mean.max():.0f} us")
            ends = cur[:, 9].astype(np.float64)                  # output index at the end of each CTA's tiles
            o = np.argsort(ends)
            outs = np.empty(ncta); outs[o] = np.diff(np.concatenate([[0.0], ends[o]]))
            ntile = cur[:, 8].astype(np.float64)
            A = np.stack([ntile, outs / 4096.0, np.ones(ncta)], axis=1)
            coef, *_ = np.linalg.lstsq(A, tiles, rcond=None)
            print(f"  outputs per CTA {outs.min():.0f}..{outs.max():.0f} (mean {outs.mean():.0f}); input tiles {ntile.min():.0f}..{ntile.max():.0f}; "
                  f"corr(duration, outputs) {np.corrcoef(tiles, outs)[0, 1]:.3f}")
            print(f"  least squares: duration = {coef[0]:.2f} us x input tiles + {coef[1]:.2f} us x (outputs / 4096) + {coef[2]:.1f} us; "
                  f"residual std {np.std(tiles - A @ coef):.1f} us")
            by_rank = {}
            for sm_id in per_sm:
                bl = sorted(b for b in range(ncta) if int(smid[b]) == sm_id)
                for r, b in enumerate(bl):
                    by_rank.setdefault(r, []).append((float(tiles[b]), float(ntile[b])))
            print("  by rank on the SM (block-index order): " + "; ".join(
                f"rank {r}: {np.mean([v[0] for v in by_rank[r]]):.0f} us for {np.mean([v[1] for v in by_rank[r]]):.1f} tiles"
                for r in sorted(by_rank)))
            order = np.argsort(tiles)
            print("  fastest CTAs (block, sm, us): " + " ".join(f"({b},{int(smid[b])},{tiles[b]:.0f})" for b in order[:8]))
            print("  slowest CTAs (block, sm, us): " + " ".join(f"({b},{int(smid[b])},{tiles[b]:.0f})" for b in order[-8:]))
            print("  by block index (us): " + " ".join(f"{tiles[b]:.0f}" for b in range(0, ncta, max(1, ncta // 37))))
            print(f"  tile-loop duration histogram (us): " + " ".join(f"{v:.0f}" for v in np.percentile(tiles, [0, 10, 50, 90, 100])))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
