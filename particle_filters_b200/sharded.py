"""Sharded SIR filter: one process per GPU, particles split into contiguous shards.

SURVEY.md section 8e.  Rank r owns the global particle indices [r N / G, (r + 1) N / G) of both the
inputs and the outputs of every step; Philox is keyed by the GLOBAL particle id and the resampling CDF
is an exact integer scan, so a G-shard run selects the same ancestors as the single-GPU run.

``exchange="mailbox"`` (default): no collective call in the step.  Every rank owns a mailbox in
peer-mapped memory (CUDA IPC); the last CTA of the producing kernel stores the rank's record straight
into the mailbox of EVERY rank over NVLink and raises a stamped flag there, the consuming kernel spins
(bounded: 20 s, then an error word is set) on the flags in its own memory:

1. partial record (max, sum e, sum e^2, weighted moments): last CTA of K1 -> ``k2_finalize_sharded``,
   which reduces the G records in rank order (identical normaliser / ESS / decision / u everywhere);
2. fixed-point weight total: last CTA of the range pre-pass -> prologue of the resampler; the exclusive
   prefix over the lower ranks is this shard's CDF offset, the sum over all ranks the CDF total the
   stratified positions are laid over;
3. the resampling kernel stores every output particle into the buffer of the rank that owns its index
   (NVLink peer stores for the few that cross a shard boundary) -- gather and cross-shard migration are
   one kernel; its last CTA publishes the post-resample moment record, whose flag also certifies that all
   migrated particles have landed -> ``k_finish_step_sharded``.

``ShardedSir.run`` / ``fused_begin`` / ``fused_step`` / ``fused_end``: the FUSED sharded step of bootstrap filters on
large shards -- ONE launch and ONE exchange per step (csrc/fused_step.cuh, sharded form; include/smc.h,
``smc_sir_step_fused_sharded``).  Every CTA's prologue finalises step t-1 and finishes step t-2 from the slot all ranks
published at the end of the launch before; a resampled quad goes through the model functor on the rank that holds its
ancestor and is stored -- particle, log-weight and fixed-point weight into the range totals -- into the memory of the
rank that owns its index; the last CTA publishes the rank's record of step t, the moments of step t-1 and its row of the
G x G table of weight contributions under one flag.  For this form the two log-weight arrays and a range-total area of
every rank are peer-mapped as well.

``exchange="nccl"``: the same kernels with three NCCL all-gathers per step instead (baseline).

No host synchronisation happens inside the step loop.  The pure-host helpers at the top
(``combine_partial_records``, ``shard_offsets``, ``migration_plan``) restate the device arithmetic for
the CPU (gloo) tests of the multi-rank logic.
"""

from __future__ import annotations

import ctypes as C
import json
import os
import time
from typing import List

import numpy as np

from . import _lib


# ------------------------------------------------------------------------------------------------
# host-side restatement of the cross-rank arithmetic (used by tests under gloo, and for planning)
# ------------------------------------------------------------------------------------------------
def combine_partial_records(records: np.ndarray) -> np.ndarray:
    """Combine (max, sum e, sum e^2, moments...) records exactly like ``combine_partials``
    in csrc/sir_kernels.cu: common max, sums rescaled by exp(m_r - M) (squared for slot 2),
    accumulated in rank order."""
    rec = np.asarray(records, dtype=np.float64)
    M = np.max(rec[:, 0])
    Mref = 0.0 if np.isneginf(M) else M
    out = np.zeros(rec.shape[1])
    out[0] = M
    for r in range(rec.shape[0]):
        sc = np.exp(rec[r, 0] - Mref) if not np.isneginf(rec[r, 0]) else 0.0
        for k in range(1, rec.shape[1]):
            out[k] += rec[r, k] * (sc * sc if k == 2 else sc)
    return out


def frac_bits_for(n_global: int, f32: bool = False) -> int:
    """Fraction bits of the fixed-point CDF (csrc/sir_kernels.cu ``frac_bits_for``): 32 for fp32
    populations, 62 - ceil(log2 N) clamped to [32, 60] for the fp64 parity mode."""
    if f32:
        return 32
    lg = 0
    while (1 << lg) < n_global:
        lg += 1
    return min(60, max(32, 62 - lg))


def shard_offsets(totals: List[int]) -> List[int]:
    """Exclusive prefix of the shards' fixed-point weight totals (rank order)."""
    out, acc = [], 0
    for t in totals:
        out.append(acc)
        acc += int(t)
    return out


def outputs_below(C_fix: int, u: float, n_global: int, f32: bool = False, total: int = 0) -> int:
    """Number of stratified positions (u + i) / N below the fixed-point CDF value C (away from
    round-off ties this equals the device's closed form).  fp32 mode lays the positions over the ACTUAL
    total of the fixed-point weights (csrc/sir_kernels.cu ``RhoMap``): C is first mapped to
    floor(C N 2^32 / total)."""
    F = frac_bits_for(n_global, f32)
    if f32 and total > 0:
        C_fix = (int(C_fix) * n_global << 32) // int(total)
    uq = int(np.floor(np.ldexp(u, F)))
    e = (C_fix + ((1 << F) - 1 - uq)) >> F
    return int(min(max(e, 0), n_global))


def migration_plan(totals: List[int], u: float, n_global: int, f32: bool = False) -> np.ndarray:
    """``plan[src, dst]`` = number of output particles whose ancestor lives on shard ``src``
    and whose index belongs to shard ``dst`` (what an alltoallv would have to move)."""
    G = len(totals)
    shard = n_global // G
    off = shard_offsets(totals) + [sum(int(t) for t in totals)]
    plan = np.zeros((G, G), dtype=np.int64)
    for s in range(G):
        lo = 0 if s == 0 else outputs_below(off[s], u, n_global, f32, off[-1])
        hi = n_global if s == G - 1 else outputs_below(off[s + 1], u, n_global, f32, off[-1])
        for d in range(G):
            a, b = max(lo, d * shard), min(hi, (d + 1) * shard)
            plan[s, d] = max(0, b - a)
    return plan


def quantised_weights_f32(logw: np.ndarray, n_global: int) -> np.ndarray:
    """Fixed-point weights of a bootstrap step in the fp32 mode, against the a-priori bound lw <= -log N
    (csrc/sir_kernels.cu ``quant_bound_for`` / ``quant_p`` / ``quant_a``): a = rni(2^(lw log2e + kshift)) with
    kshift = Fa + log2 N, Fa = min(36, 62 - ceil(log2 N)).  Python ints (the sums are exact on the device too).  The
    device evaluates 2^x with ex2.approx (2 ulp), so these integers are a restatement of the PROTOCOL (what is summed and
    exchanged), not a bit-for-bit oracle of the device's weights."""
    lg = 0
    while (1 << lg) < n_global:
        lg += 1
    fa = min(36, 62 - lg)
    kshift = np.float32(fa + np.log2(float(n_global)))
    p = np.exp2(np.float32(1.4426950408889634) * np.asarray(logw, np.float32) + kshift)
    return np.rint(p.astype(np.float64)).astype(object)


def contribution_row(a_out: np.ndarray, first_output_index: int, n_global: int, n_shards: int) -> List[int]:
    """What ONE rank publishes in its slot of the fused sharded step (csrc/fused_step.cuh, sharded form): the
    fixed-point weight of the step-t particles it PRODUCED (outputs ``first_output_index ..`` of the resampling of step
    t-1, after the functor), split by the rank that OWNS each output index."""
    shard = n_global // n_shards
    row = [0] * n_shards
    for k, a in enumerate(a_out):
        row[(first_output_index + k) // shard] += int(a)
    return row


def shard_totals_from_contributions(table) -> List[int]:
    """Column sums of the G x G contribution table ([producing rank][owning rank]) = the shards' weight totals of the
    NEXT resampling; every rank computes them alike in the prologue of the next launch (exclusive prefix over the lower
    ranks = the rank's CDF offset)."""
    G = len(table)
    return [sum(int(table[src][dst]) for src in range(G)) for dst in range(G)]


# ------------------------------------------------------------------------------------------------
# device side
# ------------------------------------------------------------------------------------------------
class _RawCuda:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(ptr), False), "version": 3}


class ShardedSir:
    """One rank of a G-way sharded SIR filter (torch.distributed must be initialised, NCCL)."""

    def __init__(self, n_global: int, nx: int, nz: int, *, dtype: str = "f32", trace_len: int = 1,
                 exchange: str = "mailbox"):
        """``exchange``: "mailbox" = records travel by NVLink peer stores issued inside the kernels
        (no collective call in the step); "nccl" = three NCCL all-gathers per step (baseline)."""
        import torch
        import torch.distributed as dist

        from .engine import SirEngine, _round4

        self.torch, self.dist = torch, dist
        self.rank, self.G = dist.get_rank(), dist.get_world_size()
        assert n_global % self.G == 0 and (n_global // self.G) % 4096 == 0, \
            "shards must be equal and tile aligned (multiple of 4096 particles)"
        self.n = n_global // self.G
        self.lib = _lib.load()
        dev = torch.device("cuda", torch.cuda.current_device())
        # particle buffers from raw cudaMalloc so that they can be exported through CUDA IPC
        elt = 4 if dtype == "f32" else 8
        ld = _round4(self.n)
        self.exchange = exchange
        self.stamp = 0
        self._raw = []
        # x0, x1, mailbox; for the fused sharded step also the two log-weight arrays and the range-total area
        sizes = [nx * ld * elt, nx * ld * elt, int(self.lib.smc_mailbox_bytes()), ld * elt, ld * elt,
                 int(self.lib.smc_shard_rsum_bytes())]
        for nbytes in sizes:
            p = C.c_void_p()
            _lib.check(self.lib.smc_dev_alloc(C.c_uint64(nbytes), C.byref(p)))
            self._raw.append(p.value)
        typestr = "<f4" if dtype == "f32" else "<f8"
        xbufs = [torch.as_tensor(_RawCuda(p, (nx, ld), typestr), device=dev) for p in self._raw[:2]]
        self.lw = [torch.as_tensor(_RawCuda(p, (ld,), typestr), device=dev) for p in self._raw[3:5]]
        self.eng = SirEngine(self.n, nx, nz, dtype=dtype, n_global=n_global, offset=self.rank * self.n,
                             trace_len=trace_len, x_buffers=xbufs, logw_buffer=self.lw[0])
        self._lw_cur = 0                 # which of the two log-weight arrays the engine's descriptor points at
        self._fused_open = False
        # exchange IPC handles; table = [x0 of every rank][x1 ...][mailboxes][logw 0 ...][logw 1 ...][range totals ...]
        mine = []
        for p in self._raw:
            h = C.create_string_buffer(64)
            _lib.check(self.lib.smc_ipc_export(C.c_void_p(p), h))
            mine.append(bytes(h.raw))
        allh = [None] * self.G
        dist.all_gather_object(allh, mine)
        nb = len(self._raw)
        table = np.zeros(nb * self.G, dtype=np.int64)
        self._opened = []
        for r in range(self.G):
            for b in range(nb):
                if r == self.rank:
                    table[b * self.G + r] = self._raw[b]
                else:
                    q = C.c_void_p()
                    _lib.check(self.lib.smc_ipc_import(allh[r][b], C.byref(q)))
                    self._opened.append(q.value)
                    table[b * self.G + r] = q.value
        self.peer_table = torch.from_numpy(table).to(dev)
        self.ctx = _lib.ShardCtx(mail_dev=self.peer_table.data_ptr() + 8 * 2 * self.G,
                                 peer_x_dev=self.peer_table.data_ptr(), rank=self.rank, n_shards=self.G, stamp=0,
                                 peer_aux_dev=self.peer_table.data_ptr() + 8 * 3 * self.G, rsum_area_dev=self._raw[5],
                                 mailbox_dev=self._raw[2])
        ps = int(self.lib.smc_partial_stride(nx))
        nm = nx + nx * (nx + 1) // 2
        self.rec_local = torch.zeros(ps, dtype=torch.float64, device=dev)
        self.rec_all = torch.zeros(self.G * ps, dtype=torch.float64, device=dev)
        self.tot_local = torch.zeros(1, dtype=torch.int64, device=dev)
        self.tot_all = torch.zeros(self.G, dtype=torch.int64, device=dev)
        self.mom_local = torch.zeros(nm, dtype=torch.float64, device=dev)
        self.mom_all = torch.zeros(self.G * nm, dtype=torch.float64, device=dev)
        dist.barrier()

    def step(self, model_id, params, thresh, z_index, seed, t, row):
        if self.exchange == "mailbox":
            return self.step_mailbox(model_id, params, thresh, z_index, seed, t, row)
        return self.step_nccl(model_id, params, thresh, z_index, seed, t, row)

    def step_mailbox(self, model_id, params, thresh, z_index, seed, t, row):
        """One sharded step = five launches, no collective: K1 (publishes the partial record),
        finalize (waits for G records), range pre-pass (publishes the weight total), K3 (waits for
        G totals, migrates by peer stores, publishes the moment record), finish (waits for G flags)."""
        eng = self.eng
        self.stamp += 1
        self.ctx.stamp = self.stamp
        kp, pp = _lib.as_doubles(params)
        _lib.check(self.lib.smc_sir_step_sharded(eng._d, C.byref(self.ctx), model_id, pp, kp.size, eng._zptr(z_index),
                                                 float(thresh), C.c_uint64(seed), t, eng._row(row), eng.stream))

    # ---- fused run: ONE launch and ONE exchange per step (bootstrap filters, fp32 SV; include/smc.h) ----------------
    def can_fuse(self, model_id, thresh) -> bool:
        n_tiles = self.n // 4096
        return bool(self.exchange == "mailbox" and model_id == _lib.SMC_MODEL_STOCH_VOL and thresh > 1.0
                    and self.eng.dtype_code == _lib.SMC_F32 and self.eng.nx == 1 and n_tiles >= 148 * 12
                    and not os.environ.get("SMC_SHARDED_NO_FUSE"))

    def fused_begin(self, model_id, params, thresh, z_index, seed, t):
        """K1 of the run's first step (publishes the partial record and the weight total)."""
        assert not self._fused_open
        eng = self.eng
        self.stamp += 1
        self.ctx.stamp = self.stamp
        kp, pp = _lib.as_doubles(params)
        _lib.check(self.lib.smc_sir_weight_sharded(eng._d, C.byref(self.ctx), model_id, pp, kp.size, eng._zptr(z_index),
                                                   float(thresh), C.c_uint64(seed), t, eng.stream))
        self._fused_open, self._fused_first = True, True

    def fused_step(self, model_id, params, thresh, z_index, seed, t, row):
        """resample(t-1) -> propagate + weigh(t) in one launch; ``row`` = statistics row of step t (rows row-1 and
        row-2 receive the header of step t-1 and the moments of step t-2 in this launch's prologue)."""
        assert self._fused_open
        eng = self.eng
        self.stamp += 1
        self.ctx.stamp = self.stamp
        kp, pp = _lib.as_doubles(params)
        out = self._lw_cur ^ 1
        _lib.check(self.lib.smc_sir_step_fused_sharded(
            eng._d, C.byref(self.ctx), out, 1 if self._fused_first else 0, model_id, pp, kp.size, eng._zptr(z_index),
            float(thresh), C.c_uint64(seed), t, None if self._fused_first else eng._row(max(row - 2, 0)),
            eng._row(max(row - 1, 0)), eng.stream))
        self._fused_first = False
        self._lw_cur = out
        eng.logw = self.lw[out]
        eng.desc.logw_dev = eng.logw.data_ptr()

    def fused_end(self, thresh, seed, t_last, row_last, resample=True):
        """Finalize of the last step, finish of the one before and (``resample``) the last step's resampling."""
        assert self._fused_open and not self._fused_first
        eng = self.eng
        self.stamp += 1
        self.ctx.stamp = self.stamp
        _lib.check(self.lib.smc_sir_fused_sharded_flush(eng._d, C.byref(self.ctx), float(thresh), C.c_uint64(seed), t_last,
                                                        eng._row(max(row_last - 1, 0)), eng._row(row_last),
                                                        1 if resample else 0, eng.stream))
        self._fused_open = False

    def run(self, model_id, params, T, thresh, seed, t0=0, z0=0, row0=0, fused=None):
        """T steps (observations z0.., statistics rows row0.., Philox steps t0..); no host synchronisation."""
        if fused is None:
            fused = self.can_fuse(model_id, thresh) and T >= 2
        if not fused:
            for k in range(T):
                self.step(model_id, params, thresh, z0 + k, seed, t0 + k, row0 + k)
            return False
        self.fused_begin(model_id, params, thresh, z0, seed, t0)
        for k in range(1, T):
            self.fused_step(model_id, params, thresh, z0 + k, seed, t0 + k, row0 + k)
        self.fused_end(thresh, seed, t0 + T - 1, row0 + T - 1)
        return True

    def step_nccl(self, model_id, params, thresh, z_index, seed, t, row):
        eng, lib, dist = self.eng, self.lib, self.dist
        eng.propagate_weight(model_id, params, _lib.SMC_MODE_STEP, z_index=z_index, seed=seed, t=t)
        _lib.check(lib.smc_sir_reduce_partials(eng._d, C.c_void_p(self.rec_local.data_ptr()), eng.stream))
        dist.all_gather_into_tensor(self.rec_all, self.rec_local)
        eng.finalize(thresh, u=-1.0, seed=seed, t=t, row=row, parts=self.rec_all, n_parts=self.G)
        _lib.check(lib.smc_sir_shard_weight_total(eng._d, C.c_void_p(self.tot_local.data_ptr()), eng.stream))
        dist.all_gather_into_tensor(self.tot_all, self.tot_local)
        _lib.check(lib.smc_sir_resample_sharded(eng._d, C.c_void_p(self.tot_all.data_ptr()), self.rank, self.G,
                                                C.c_void_p(self.peer_table.data_ptr()), eng.stream))
        _lib.check(lib.smc_sir_reduce_moment_partials(eng._d, C.c_void_p(self.mom_local.data_ptr()), eng.stream))
        dist.all_gather_into_tensor(self.mom_all, self.mom_local)     # also orders the peer stores
        eng.finish_step(row=row, mom_parts=self.mom_all, n_mom_parts=self.G)

    def check(self):
        """Raise if a flag wait on this rank timed out (a peer never published: it failed, stepped a different
        number of times, or died).  Synchronises the stream."""
        if self.exchange != "mailbox":
            return
        dev = self.torch.device("cuda", self.torch.cuda.current_device())
        off = int(self.lib.smc_mailbox_error_offset())
        word = self.torch.as_tensor(_RawCuda(self._raw[2] + off, (1,), "<i8"), device=dev)
        v = int(word.cpu().numpy().astype(np.uint64)[0])
        if v:
            what = {1: "partial record", 2: "weight total", 3: "moment record", 4: "fused-step slot"}.get((v >> 4) & 0xF, "?")
            raise _lib.SmcError(f"rank {self.rank}: timed out waiting for the {what} of rank {v & 0xF} "
                                f"(stamp {(v >> 8) & ((1 << 55) - 1)}); the sharded filter state is invalid")

    def shard_totals(self) -> np.ndarray:
        """The G fixed-point weight totals of the last resampling step (rank order), as Python ints."""
        if self.exchange == "nccl":
            return self.tot_all.cpu().numpy().astype(object)
        dev = self.torch.device("cuda", self.torch.cuda.current_device())
        mail = self.torch.as_tensor(_RawCuda(self._raw[2] + 8 * 3 * 16, (16,), "<i8"), device=dev)   # Mailbox.tot
        return mail[: self.G].cpu().numpy().astype(np.uint64).astype(object)

    def close(self):
        self.torch.cuda.synchronize()
        self.dist.barrier()
        for q in self._opened:
            self.lib.smc_ipc_close(C.c_void_p(q))
        self.dist.barrier()
        for p in self._raw:
            self.lib.smc_dev_free(C.c_void_p(p))


def init_distributed():
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    return dist


def run_sharded_bench(args, sv_observations, measured_peak, ClockSampler, METRIC, UNIT):
    """bench.py's N > 1 arm: BASELINE config 3 (SV, N = 2^26 total) sharded over the ranks (strong scaling),
    preceded by a parity self-check against the single-GPU engine on rank 0 and followed by a weak-scaling
    sub-record (2^26 particles PER GPU)."""
    import torch

    from .engine import SirEngine
    from .ssm import StochVolSSM

    dist = init_distributed()
    rank, G = dist.get_rank(), dist.get_world_size()
    N = 1 << args.log2n
    K, W = args.steps, max(3, args.warmup)
    ALPHA, SIGMA, BETA = 0.95, 0.2, 1.0
    Q, R = np.array([[SIGMA ** 2]]), np.array([[BETA ** 2]])
    mid, params = StochVolSSM(ALPHA, BETA).device_functor(Q, R)
    Y = sv_observations(W + K + 2)
    thresh, seed = 2.0, 42
    L0 = np.array([[SIGMA / np.sqrt(1 - ALPHA ** 2)]])

    def timed_run(n_total, k_steps):
        sh = ShardedSir(n_total, 1, 1, dtype="f32", trace_len=W + k_steps + 2, exchange=getattr(args, "exchange", "mailbox"))
        sh.eng.init_gaussian(np.zeros(1), L0, seed)
        sh.eng.set_observations(Y.reshape(-1, 1))
        sh.run(mid, params, W, thresh, seed)           # fused (one launch, one exchange per step) where it applies
        torch.cuda.synchronize()
        return sh

    def measure(sh, n_steps, spin):
        """Steady state on an initialised filter: `spin` untimed steps (clock warm-up; a FIXED number -- every rank
        must issue the same sequence of stamps), then n_steps timed ones between CUDA events.  Fused: one launch per
        step inside an open run (K1 before, flush + last resampling after, both outside the timed region).
        Statistics rows of the timed steps: W+2 .. W+n_steps+1.  Returns (ms max over ranks, launches, rows)."""
        fz = sh.can_fuse(mid, thresh)
        zmax = Y.shape[0] - 1

        def go(k0, cnt, row0=None, row=None):
            for k in range(k0, k0 + cnt):
                r = row if row is not None else row0 + (k - k0)
                if fz:
                    sh.fused_step(mid, params, thresh, min(k, zmax), seed, k, r)
                else:
                    sh.step(mid, params, thresh, min(k, zmax), seed, k, r)

        if fz:
            sh.fused_begin(mid, params, thresh, W, seed, W)
        for kk in range(0, spin, 250):
            go(W + 1 + kk, min(250, spin - kk), row=W + 2)
            torch.cuda.synchronize()
        flag = torch.zeros(1, device="cuda")
        dist.all_reduce(flag)                       # ranks leave the spin at different times: line them up
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lib = _lib.load()
        launches0 = int(lib.smc_launch_count())
        e0.record()
        go(W + 1 + spin, n_steps, row0=W + 2)
        e1.record()
        torch.cuda.synchronize()
        launches = int(lib.smc_launch_count()) - launches0
        if fz:
            sh.fused_end(thresh, seed, W + spin + n_steps, W + n_steps + 1)
            torch.cuda.synchronize()
        dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        rows = sh.eng.trace[W + 2: W + n_steps + 2].cpu().numpy()
        return float(ms.item()), launches, rows, fz

    # ---- parity: the first W statistic rows and the particle set after W steps vs ONE GPU, same N and seed ----
    sh = timed_run(N, K)
    rows_w = sh.eng.trace[:W].cpu().numpy()
    local = sh.eng.get_particles()[:, 0]
    gathered = [None] * G
    dist.all_gather_object(gathered, local)
    parity = None
    if rank == 0:
        ref = SirEngine(N, 1, 1, dtype="f32", trace_len=W)
        ref.init_gaussian(np.zeros(1), L0, seed)
        ref.set_observations(Y[:W].reshape(-1, 1))
        ref.enqueue(mid, params, W, thresh, seed, fused=False)
        rrows = ref.trace[:W].cpu().numpy()
        xref = ref.get_particles()[:, 0]
        xsh = np.concatenate(gathered)
        parity = {"steps_compared": W, "reference": "single-GPU three-launch engine, same N and seed, on rank 0",
                  "sharded_pipeline": ("fused: one launch + one exchange per step" if sh.can_fuse(mid, thresh)
                                       else "three / five launches per step"),
                  "identical_particle_fraction": float(np.mean(xsh == xref)),
                  "decisions_equal": bool(np.array_equal(rows_w[:, 2], rrows[:, 2])),
                  "uniforms_equal": bool(np.array_equal(rows_w[:, 3], rrows[:, 3])),
                  "max_abs_log_norm_diff": float(np.max(np.abs(rows_w[:, 0] - rrows[:, 0]))),
                  "max_rel_ess_diff": float(np.max(np.abs(rows_w[:, 1] - rrows[:, 1]) / rrows[:, 1])),
                  "max_abs_mean_diff": float(np.max(np.abs(rows_w[:, 4] - rrows[:, 4])))}
        parity["pass"] = bool(parity["decisions_equal"] and parity["uniforms_equal"] and
                              parity["identical_particle_fraction"] > 0.999 and parity["max_abs_log_norm_diff"] < 1e-6
                              and parity["max_rel_ess_diff"] < 1e-6 and parity["max_abs_mean_diff"] < 1e-6)
        parity["tolerances"] = "particles bit-identical (> 99.9 %), |d log-norm| < 1e-6, rel dESS < 1e-6 (fp32 block sums over "                                "another partition), |d mean| < 1e-6"
        del ref, xref, xsh
        torch.cuda.empty_cache()
    del gathered, local

    clocks = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    if rank == 0:
        clocks.start()
    # keep every GPU on the workload while the sampler takes its first readings (outside the timed region)
    ms_total, launches, rows, fused = measure(sh, K, 4000)
    ms_step = ms_total / K
    # end to end: host observation in, host moments out, every step, through the same engine
    Ke = max(3, min(K, 20))
    zpin = torch.zeros(1, dtype=torch.float64).pin_memory()
    dist.barrier()
    t0 = time.perf_counter()
    for k in range(Ke):
        zpin[0] = float(Y[k])
        sh.eng.zbuf[0].copy_(zpin, non_blocking=True)
        sh.step(mid, params, thresh, 0, seed, 1000 + k, 0)
        _ = sh.eng.read_row(0)[4]
    torch.cuda.synchronize()
    dist.barrier()
    e2e_dt = torch.tensor([(time.perf_counter() - t0) / Ke], dtype=torch.float64, device="cuda")
    dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    clk = clocks.stop() if rank == 0 else None
    sh.check()
    ok = bool(np.all(np.isfinite(rows[:, :5])) and np.all(rows[:, 2] == 1.0))
    sh.close()

    # ---- weak scaling: 2^log2n particles PER GPU ----------------------------------------------------------
    weak = None
    if not getattr(args, "no_weak", False) and N * G <= (1 << 30):
        Kw = min(K, 50)
        shw = timed_run(N * G, Kw)
        msw, _, wrows, _ = measure(shw, Kw, 0)
        weak = {"n_particles_total": N * G, "n_particles_per_gpu": N, "steps": Kw, "ms_per_step": msw / Kw,
                "value": N * G / (msw / Kw * 1e-3), "unit": UNIT,
                "finite": bool(np.all(np.isfinite(wrows[:, :5])))}
        shw.close()

    peak, peak_src = measured_peak()
    value = N / (ms_step * 1e-3)
    # fused: one launch, 16 B per particle-update (R logw, R x, W x, W logw); otherwise the three-launch form: K1 12 B
    # + resampler 12 B (+ 4 B range pre-pass on steps that need it)
    step_bytes = 16.0 if fused else 24.0
    if rank == 0:
        assert ok, "non-finite statistics in the sharded run"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": G, "steps": K, "warmup": W,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"SV SIR bootstrap filter N=2^{args.log2n} total, {G} shards of {N // G}, "
                                   "systematic resampling every step with cross-shard migration "
                                   "(BASELINE.json configs[2])",
                       "n_particles": N, "nx": 1, "nz": 1, "parallelism": f"particle-sharded x{G}",
                       "exchange": (("ONE launch + ONE exchange per step: fused resample -> propagate kernel on every shard, "
                                     "records / migrating particles / range totals by in-kernel NVLink peer stores and "
                                     "atomics, stamped flags (no collective call)") if fused else
                                    "in-kernel NVLink peer stores + flags (no collective call)" if sh.exchange == "mailbox"
                                    else "3 NCCL all-gathers per step"),
                       "l2": "per-step working set 3 x N/G x 4 B; no flush (strong scaling shrinks it)"},
            "clocks": clk,
            "e2e": {"value": N / float(e2e_dt.item()), "unit": UNIT, "h2d_bytes_per_step": 8,
                    "d2h_bytes_per_step": 8 * int(sh.eng.stride), "ms_per_step": float(e2e_dt.item()) * 1e3},
            "gpu_launches": launches, "parity": parity, "weak_scaling": weak,
            "roofline": {"bound": "hbm", "kernel": "whole sharded step (per-kernel split is taken at N=1)",
                         "achieved": step_bytes * value / 1e9 / G, "peak": peak, "unit": "GB/s",
                         "frac": step_bytes * value / 1e9 / G / peak, "traffic": None, "peak_source": peak_src,
                         "bytes_moved_per_particle_update": step_bytes},
            "roofline_step": {"bytes_moved_per_particle_update": step_bytes, "achieved": step_bytes * value / 1e9,
                              "peak": peak * G, "unit": "GB/s", "frac": step_bytes * value / 1e9 / (peak * G),
                              "contract_bytes_per_particle_update": 32.0},
            "cpu_baseline": None,
            "filter": {"mean_last": float(rows[-1, 4]), "ess_last": float(rows[-1, 1])},
        }
        print(json.dumps(line))
    dist.destroy_process_group()
