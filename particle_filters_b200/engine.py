"""Device-resident SIR engine: torch tensors hold the particle arrays, every
operation is a libsmc (CUDA, sm_100a) call enqueued on the current stream.

PyTorch is plumbing here (allocation, streams, H2D/D2H); no torch op touches
the particle data on the hot path.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import SirDesc, check


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise _lib.SmcError("the SMC engine needs a CUDA device (no CPU fallback)")
    return torch


def _round4(n: int) -> int:
    return (n + 3) // 4 * 4


class SirEngine:
    """One shard of an SIR particle population on one GPU.

    Arrays (SoA, see include/smc.h): ``x0/x1`` ping-pong particles ``(nx, ld)``,
    ``logw`` ``(ld,)``, optional int32 ``ancestors``, a byte workspace, and a
    device trace of per-step statistic rows.
    """

    def __init__(self, n: int, nx: int, nz: int, *, dtype: str = "f32", device=None,
                 n_global: Optional[int] = None, offset: int = 0, write_ancestors: bool = False,
                 generic_moments: bool = False, post_moments_generic: bool = False, trace_len: int = 1,
                 x_buffers=None, logw_buffer=None):
        torch = _torch()
        self.lib = _lib.load()
        self.torch = torch
        cur_dev = torch.device("cuda", torch.cuda.current_device())
        self.device = cur_dev if device is None else torch.device(device)
        if self.device.type != "cuda" or (self.device.index is not None and self.device.index != cur_dev.index):
            # libsmc launches on the CURRENT device and keeps per-process function attributes: one process per GPU
            # (torchrun), select the GPU with torch.cuda.set_device() before building engines
            raise ValueError(f"SirEngine runs on the process's current CUDA device ({cur_dev}); got device={device!r}. "
                             "Call torch.cuda.set_device() first (one process per GPU).")
        self.device = cur_dev
        self.n, self.nx, self.nz = int(n), int(nx), int(nz)
        self.n_global = self.n if n_global is None else int(n_global)
        self.offset = int(offset)
        self.dtype_code = {"f32": _lib.SMC_F32, "f64": _lib.SMC_F64}[dtype]
        self.tdtype = torch.float32 if dtype == "f32" else torch.float64
        self.npdtype = np.float32 if dtype == "f32" else np.float64
        self.ld = _round4(self.n)
        self.generic_moments = bool(generic_moments) or self.nx > 4
        with torch.cuda.device(self.device):
            if x_buffers is not None:      # caller-owned (e.g. IPC-shareable) ping-pong buffers
                self.x = list(x_buffers)
                assert all(tuple(b.shape) == (self.nx, self.ld) and b.dtype == self.tdtype for b in self.x)
            else:
                self.x = [torch.zeros((self.nx, self.ld), dtype=self.tdtype, device=self.device) for _ in range(2)]
            if logw_buffer is not None:    # caller-owned (peer-mapped) log-weight array
                assert tuple(logw_buffer.shape) == (self.ld,) and logw_buffer.dtype == self.tdtype
                self.logw = logw_buffer
            else:
                self.logw = torch.zeros((self.ld,), dtype=self.tdtype, device=self.device)
            need_anc = write_ancestors or self.nx > 4
            self.ancestors = (torch.zeros((self.ld,), dtype=torch.int32, device=self.device)
                              if need_anc else None)
            ws_bytes = int(self.lib.smc_sir_workspace_bytes(self.n, self.nx, self.dtype_code))
            self.ws = torch.zeros((ws_bytes,), dtype=torch.uint8, device=self.device)
            self.stride = _lib.stats_stride(self.nx)
            self.trace = torch.zeros((max(1, trace_len), self.stride), dtype=torch.float64, device=self.device)
            self.zbuf = torch.zeros((max(1, trace_len), self.nz), dtype=torch.float64, device=self.device)
            # pinned staging for the per-step host <-> device traffic of the class API (one observation in,
            # one statistics row out): asynchronous copies on the engine's stream, one stream sync per step
            self._z_pin = torch.zeros((self.nz,), dtype=torch.float64).pin_memory()
            self._row_pin = torch.zeros((self.stride,), dtype=torch.float64).pin_memory()
        flags = 0
        if write_ancestors:
            flags |= _lib.SMC_FLAG_WRITE_ANCESTORS
        if self.generic_moments:
            flags |= _lib.SMC_FLAG_GENERIC_MOMENTS
        if post_moments_generic:
            flags |= _lib.SMC_FLAG_POST_MOMENTS_GENERIC
        self.desc = SirDesc(
            n=self.n, n_global=self.n_global, offset=self.offset, ld=self.ld, nx=self.nx, nz=self.nz,
            dtype=self.dtype_code, flags=flags, x0_dev=self.x[0].data_ptr(), x1_dev=self.x[1].data_ptr(),
            logw_dev=self.logw.data_ptr(),
            ancestors_dev=self.ancestors.data_ptr() if self.ancestors is not None else None,
            ws_dev=self.ws.data_ptr(), ws_bytes=ws_bytes)
        self._d = C.byref(self.desc)
        self.reset()

    # ------------------------------------------------------------------ plumbing
    @property
    def stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _row(self, t: int = 0):
        return C.c_void_p(self.trace.data_ptr() + 8 * self.stride * (t % self.trace.shape[0]))

    def sync(self):
        self.torch.cuda.current_stream(self.device).synchronize()

    # ------------------------------------------------------------------ state
    def reset(self, shift=None):
        if shift is None:
            check(self.lib.smc_sir_reset(self._d, None, self.stream))
        else:
            keep, p = _lib.as_doubles(shift)
            check(self.lib.smc_sir_reset(self._d, p, self.stream))

    def current_buffer(self) -> int:
        cur = C.c_int32(0)
        check(self.lib.smc_sir_current_buffer(self._d, C.byref(cur), self.stream))
        return int(cur.value)

    def set_particles(self, particles: np.ndarray, shift=None):
        """Upload host particles ``(n, nx)`` into buffer 0 and reset the control block
        (weights uniform)."""
        torch = self.torch
        p = np.ascontiguousarray(np.asarray(particles, dtype=np.float64).reshape(self.n, self.nx).T)
        host = torch.from_numpy(p.astype(self.npdtype, copy=False))
        self.x[0][:, : self.n].copy_(host, non_blocking=False)
        self.reset(shift)

    def upload_particles(self, particles: np.ndarray):
        """Overwrite the CURRENT particle buffer with host particles ``(n, nx)``; weights, normaliser and the rest of
        the control block stay as they are (assigning ``state.particles`` in the reference leaves the weights alone)."""
        torch = self.torch
        p = np.ascontiguousarray(np.asarray(particles, dtype=np.float64).reshape(self.n, self.nx).T)
        self.x[self.current_buffer()][:, : self.n].copy_(torch.from_numpy(p.astype(self.npdtype, copy=False)))

    def get_particles(self) -> np.ndarray:
        """Current particles as host float64 ``(n, nx)``."""
        cur = self.current_buffer()
        return self.x[cur][:, : self.n].to("cpu", dtype=self.torch.float64).numpy().T.copy()

    def init_gaussian(self, mean, Lc, seed: int):
        km, pm = _lib.as_doubles(mean)
        kl, pl = _lib.as_doubles(Lc)
        self.reset(mean)
        check(self.lib.smc_sir_init_gaussian(self._d, pm, pl, C.c_uint64(seed), self.stream))

    def weights(self) -> np.ndarray:
        w = self.torch.empty((self.n,), dtype=self.torch.float64, device=self.device)
        check(self.lib.smc_sir_export_weights(self._d, C.c_void_p(w.data_ptr()), self.stream))
        return w.cpu().numpy()

    def upload_soa(self, a: np.ndarray, rows: int):
        """Host ``(n, rows)`` -> device SoA ``(rows, ld)`` in the engine dtype."""
        torch = self.torch
        t = torch.zeros((rows, self.ld), dtype=self.tdtype, device=self.device)
        src = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(self.n, rows).T)
        t[:, : self.n].copy_(torch.from_numpy(src.astype(self.npdtype, copy=False)))
        return t

    def set_observations(self, Z: np.ndarray):
        """Stage a block of observations ``(T, nz)`` on the device (fp64)."""
        torch = self.torch
        Z = np.ascontiguousarray(np.asarray(Z, dtype=np.float64).reshape(-1, self.nz))
        if Z.shape[0] > self.zbuf.shape[0]:
            self.zbuf = torch.zeros(Z.shape, dtype=torch.float64, device=self.device)
        self.zbuf[: Z.shape[0]].copy_(torch.from_numpy(Z))
        return Z.shape[0]

    def set_observation(self, z: np.ndarray, index: int = 0):
        """Stage ONE observation (nz,) into slot ``index`` through pinned memory (no sync)."""
        self._z_pin.numpy()[:] = np.asarray(z, dtype=np.float64).reshape(-1)
        self.zbuf[index].copy_(self._z_pin, non_blocking=True)

    def _zptr(self, t: int):
        return C.c_void_p(self.zbuf.data_ptr() + 8 * self.nz * t)

    # ------------------------------------------------------------------ kernels
    def propagate_weight(self, model_id: int, params, mode: int, z_index: int = 0, noise=None,
                         seed: int = 0, t: int = 0):
        kp, pp = _lib.as_doubles(params)
        check(self.lib.smc_sir_propagate_weight(
            self._d, model_id, pp, kp.size, mode, self._zptr(z_index),
            C.c_void_p(noise.data_ptr()) if noise is not None else None,
            C.c_uint64(seed), t, self.stream))

    def finalize(self, thresh: float, u: float = -1.0, seed: int = 0, t: int = 0, row: int = 0,
                 parts=None, n_parts: int = 0):
        check(self.lib.smc_sir_finalize(
            self._d, float(thresh), float(u), C.c_uint64(seed), t, self._row(row),
            C.c_void_p(parts.data_ptr()) if parts is not None else None, n_parts, self.stream))

    def set_uniform(self, u: float):
        check(self.lib.smc_sir_set_uniform(self._d, float(u), self.stream))

    def resample(self, row: int = 0, cdf_offset: int = 0, peers=None, n_shards: int = 0):
        check(self.lib.smc_sir_resample(
            self._d, C.c_uint64(cdf_offset), self._row(row),
            C.c_void_p(peers.data_ptr()) if peers is not None else None, n_shards, self.stream))

    def jitter(self, Lq: np.ndarray, scale: float, seed: int, t: int):
        """On-chip post-resample regularisation (no-op on the device when the step did not resample)."""
        kl, pl = _lib.as_doubles(Lq)
        check(self.lib.smc_sir_jitter(self._d, pl, float(scale), C.c_uint64(seed), t, self.stream))

    def resample_multinomial(self, u: Optional[np.ndarray] = None, seed: int = 0, t: int = 0):
        """Multinomial resampling of the current weighted set, all on the device: export the fp64
        weights, draw ancestors (``u`` = the host's ``rng.random(n)`` in reference-noise mode, else
        on-chip Philox), gather into the other ping-pong buffer.  ``finish_step`` flips the buffers."""
        torch = self.torch
        assert self.n == self.n_global, "multinomial resampling is single-shard"
        if getattr(self, "_mn_ws", None) is None:
            self._mn_w = torch.empty((self.n,), dtype=torch.float64, device=self.device)
            self._mn_bytes = int(self.lib.smc_multinomial_workspace_bytes(self.n))
            self._mn_ws = torch.empty((self._mn_bytes,), dtype=torch.uint8, device=self.device)
            self._mn_anc = self.ancestors if self.ancestors is not None else \
                torch.empty((self.n,), dtype=torch.int32, device=self.device)
        ud = None
        if u is not None:
            ud = torch.from_numpy(np.ascontiguousarray(u, dtype=np.float64).reshape(-1)).to(self.device)
            assert ud.numel() == self.n
        check(self.lib.smc_sir_export_weights(self._d, C.c_void_p(self._mn_w.data_ptr()), self.stream))
        check(self.lib.smc_multinomial_ancestors(
            C.c_void_p(self._mn_w.data_ptr()), self.n, self.n, C.c_void_p(ud.data_ptr()) if ud is not None else None,
            C.c_uint64(seed), int(t), C.c_void_p(self._mn_anc.data_ptr()), C.c_void_p(self._mn_ws.data_ptr()),
            C.c_uint64(self._mn_bytes), self.stream))
        cur = self.current_buffer()
        check(self.lib.smc_gather(C.c_void_p(self.x[cur].data_ptr()), C.c_void_p(self.x[cur ^ 1].data_ptr()),
                                  C.c_void_p(self._mn_anc.data_ptr()), self.n, self.ld, self.nx, self.dtype_code,
                                  self.stream))

    def finish_step(self, row: int = 0, mom_parts=None, n_mom_parts: int = 0):
        check(self.lib.smc_sir_finish_step(
            self._d, self._row(row),
            C.c_void_p(mom_parts.data_ptr()) if mom_parts is not None else None, n_mom_parts, self.stream))

    def step(self, model_id: int, params, thresh: float, z_index: int = 0, noise=None, u: float = -1.0,
             seed: int = 0, t: int = 0, row: int = 0):
        kp, pp = _lib.as_doubles(params)
        check(self.lib.smc_sir_step(
            self._d, model_id, pp, kp.size, self._zptr(z_index),
            C.c_void_p(noise.data_ptr()) if noise is not None else None,
            float(thresh), float(u), C.c_uint64(seed), t, self._row(row), self.stream))

    def step_host(self, model_id: int, params_c, z: np.ndarray, thresh: float, seed: int = 0, t: int = 0,
                  row: int = 0) -> np.ndarray:
        """One whole step with the observation coming from and the statistics row going to the host, in a single
        library call (pinned staging on both sides, one stream synchronisation).  ``params_c`` = the
        ``_lib.as_doubles(params)`` pair, kept by the caller across steps."""
        kp, pp = params_c
        self._z_pin.numpy()[:] = np.asarray(z, dtype=np.float64).reshape(-1)
        check(self.lib.smc_sir_step_host(
            self._d, model_id, pp, kp.size, self._zptr(0), C.c_void_p(self._z_pin.data_ptr()), float(thresh), -1.0,
            C.c_uint64(seed), t, self._row(row), C.c_void_p(self._row_pin.data_ptr()), int(self.stride), self.stream))
        return self._row_pin.numpy().copy()

    def step_weight(self, model_id: int, params, thresh: float, z_index: int = 0, noise=None, u: float = -1.0,
                    seed: int = 0, t: int = 0, row: int = 0):
        """First half of ``step``: K1 with the finalize folded into its last CTA (one launch)."""
        kp, pp = _lib.as_doubles(params)
        check(self.lib.smc_sir_step_weight(
            self._d, model_id, pp, kp.size, self._zptr(z_index),
            C.c_void_p(noise.data_ptr()) if noise is not None else None,
            float(thresh), float(u), C.c_uint64(seed), t, self._row(row), self.stream))

    def step_resample(self, phases: int = 3, row: int = 0):
        """Second half of ``step``: 1 = range pre-pass, 2 = K3 with the fused finish, 3 = both."""
        check(self.lib.smc_sir_step_resample(self._d, int(phases), self._row(row), self.stream))

    def read_row(self, row: int = 0) -> np.ndarray:
        """Statistics row -> host (pinned staging, one stream synchronisation)."""
        self._row_pin.copy_(self.trace[row % self.trace.shape[0]], non_blocking=True)
        self.torch.cuda.current_stream(self.device).synchronize()
        return self._row_pin.numpy().copy()

    def run(self, model_id: int, params, Z: np.ndarray, thresh: float, seed: int, t0: int = 0,
            fused: Optional[bool] = None) -> np.ndarray:
        """Enqueue ``len(Z)`` SIR steps without any host synchronisation and return
        the statistic rows ``(T, stats_stride)`` (one D2H at the end).

        Filters that resample every step (``thresh > 1``) take the fused pipeline when the
        configuration allows it (``fused=None``): the resampler of step t-1 hands its output
        straight to the propagate + likelihood functor of step t (two launches per step,
        20 B per particle-update); ``fused=False`` forces the three-launch step."""
        torch = self.torch
        T = self.set_observations(Z)
        if self.trace.shape[0] < T:
            self.trace = torch.zeros((T, self.stride), dtype=torch.float64, device=self.device)
        self.enqueue(model_id, params, T, thresh, seed, t0=t0, fused=fused)
        return self.trace[:T].cpu().numpy()

    def enqueue(self, model_id: int, params, T: int, thresh: float, seed: int, t0: int = 0, z0: int = 0,
                row0: int = 0, fused: Optional[bool] = None):
        """Enqueue T steps (observations ``z0 .. z0+T-1`` of the staged block, statistics rows
        ``row0 .. row0+T-1``, Philox steps ``t0 ..``) on the current stream; no synchronisation."""
        if fused is None:
            fused = self.can_fuse(model_id, thresh) and T >= 2
        if fused:
            self._run_fused(model_id, params, T, thresh, seed, t0, z0, row0)
        else:
            for k in range(T):
                self.step(model_id, params, thresh, z_index=z0 + k, seed=seed, t=t0 + k, row=row0 + k)
        return fused

    def can_fuse(self, model_id: int, thresh: float) -> bool:
        """Whether ``smc_sir_step_fused`` serves this configuration (include/smc.h)."""
        ok_model = (model_id == _lib.SMC_MODEL_STOCH_VOL and self.nx == 1) or \
            (model_id == _lib.SMC_MODEL_LINEAR_GAUSSIAN and self.nx == 2 and self.nz == 2)
        n_tiles = -(-self.n // 4096)
        return bool(ok_model and thresh > 1.0 and self.dtype_code == _lib.SMC_F32 and self.n == self.n_global
                    and self.offset == 0 and self.ancestors is None and not self.generic_moments
                    and not (self.desc.flags & _lib.SMC_FLAG_POST_MOMENTS_GENERIC) and n_tiles >= 148 * 12)

    def _run_fused(self, model_id, params, T, thresh, seed, t0, z0=0, row0=0, events=None):
        """``events``: optional callable(tag) invoked between launch groups (bench.py's per-kernel timing)."""
        torch = self.torch
        if getattr(self, "logw2", None) is None:
            self.logw2 = torch.zeros_like(self.logw)
        kp, pp = _lib.as_doubles(params)
        cur, alt = self.logw, self.logw2
        desc = SirDesc.from_buffer_copy(self.desc)
        dref = C.byref(desc)
        check(self.lib.smc_sir_step_weight(dref, model_id, pp, kp.size, self._zptr(z0), None, float(thresh), -1.0,
                                           C.c_uint64(seed), t0, self._row(row0), self.stream))
        check(self.lib.smc_sir_step_resample(dref, 1, self._row(row0), self.stream))
        for k in range(1, T):
            desc.logw_dev = cur.data_ptr()
            for ph in ((1, 2) if events else (3,)):
                if events:
                    events("fused" if ph == 1 else "prepass")
                check(self.lib.smc_sir_step_fused(dref, C.c_void_p(alt.data_ptr()), model_id, pp, kp.size,
                                                  self._zptr(z0 + k), float(thresh), C.c_uint64(seed), t0 + k,
                                                  self._row(row0 + k - 1), self._row(row0 + k), ph, self.stream))
            if events:
                events("end")
            cur, alt = alt, cur
        desc.logw_dev = cur.data_ptr()
        check(self.lib.smc_sir_step_resample(dref, 2, self._row(row0 + T - 1), self.stream))
        # the engine's canonical log-weight array is the one written last
        self.logw, self.logw2 = cur, alt
        self.desc.logw_dev = self.logw.data_ptr()

    # ------------------------------------------------------------------ bridge (host callbacks)
    def bridge_predict(self, gx: Optional[np.ndarray], noise: np.ndarray, Lq: np.ndarray, scale: float = 1.0,
                       target_other: bool = False):
        """``x = gx + noise @ (scale Lq)^T``; ``gx=None`` with ``target_other`` jitters the
        resampling destination in place (particle_filter.py:212-218)."""
        g = self.upload_soa(gx, self.nx) if gx is not None else None
        nz = self.upload_soa(noise, self.nx)
        kl, pl = _lib.as_doubles(Lq)
        check(self.lib.smc_bridge_predict(self._d, C.c_void_p(g.data_ptr()) if g is not None else None,
                                          C.c_void_p(nz.data_ptr()), pl, float(scale), int(target_other),
                                          self.stream))

    def bridge_loglik(self, zpred: np.ndarray, z: np.ndarray, LRinv: np.ndarray, wprev: Optional[np.ndarray]):
        torch = self.torch
        zp = self.upload_soa(zpred, self.nz)
        kz, pz = _lib.as_doubles(z)
        kl, pl = _lib.as_doubles(LRinv)
        wp = None
        if wprev is not None:
            wp = torch.zeros((self.ld,), dtype=self.tdtype, device=self.device)
            wp[: self.n].copy_(torch.from_numpy(np.asarray(wprev, dtype=self.npdtype)))
        check(self.lib.smc_bridge_loglik(self._d, C.c_void_p(zp.data_ptr()), pz, pl,
                                         C.c_void_p(wp.data_ptr()) if wp is not None else None, self.stream))


# ---------------------------------------------------------------------- stand-alone helpers
def systematic_ancestors(weights: np.ndarray, u: float, device=None) -> np.ndarray:
    """GPU systematic resampling of normalised fp64 weights -> int64 ancestors
    (reference ``ParticleFilter._systematic_resample``, particle_filter.py:146-171)."""
    torch = _torch()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    w = np.ascontiguousarray(weights, dtype=np.float64).reshape(-1)
    n = w.shape[0]
    ld = _round4(n)
    wd = torch.zeros((ld,), dtype=torch.float64, device=dev)
    wd[:n].copy_(torch.from_numpy(w))
    anc = torch.zeros((ld,), dtype=torch.int32, device=dev)
    ws_bytes = int(lib.smc_sir_workspace_bytes(n, 1, _lib.SMC_F64))
    ws = torch.zeros((ws_bytes,), dtype=torch.uint8, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    check(lib.smc_systematic_ancestors(C.c_void_p(wd.data_ptr()), n, float(u), C.c_void_p(anc.data_ptr()),
                                       C.c_void_p(ws.data_ptr()), C.c_uint64(ws_bytes), st))
    return anc[:n].cpu().numpy().astype(np.int64)


def multinomial_ancestors(weights: np.ndarray, u: Optional[np.ndarray] = None, size: Optional[int] = None,
                          seed: int = 0, t: int = 0, device=None) -> np.ndarray:
    """GPU multinomial resampling of fp64 probabilities -> int64 ancestors (reference
    ``ParticleFilter._multinomial_resample``, particle_filter.py:173-186, i.e. ``rng.choice(N, size, p=w)``).
    ``u``: the uniforms ``rng.random(size)`` that ``choice`` would have drawn (reference-noise mode);
    ``None`` draws them on chip from Philox ``(seed, t)``."""
    torch = _torch()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    w = np.ascontiguousarray(weights, dtype=np.float64).reshape(-1)
    n = w.shape[0]
    n_out = int(size) if size is not None else (n if u is None else int(np.size(u)))
    wd = torch.from_numpy(w).to(dev)
    ud = None
    if u is not None:
        ud = torch.from_numpy(np.ascontiguousarray(u, dtype=np.float64).reshape(-1)).to(dev)
        assert ud.numel() == n_out
    anc = torch.empty((n_out,), dtype=torch.int32, device=dev)
    ws_bytes = int(lib.smc_multinomial_workspace_bytes(n))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    check(lib.smc_multinomial_ancestors(C.c_void_p(wd.data_ptr()), n, n_out,
                                        C.c_void_p(ud.data_ptr()) if ud is not None else None,
                                        C.c_uint64(seed), int(t), C.c_void_p(anc.data_ptr()),
                                        C.c_void_p(ws.data_ptr()), C.c_uint64(ws_bytes), st))
    return anc.cpu().numpy().astype(np.int64)


def gather_rows(particles: np.ndarray, idx: np.ndarray, device=None) -> np.ndarray:
    """``particles[idx]`` on the GPU (reference particle_filter.py:209)."""
    torch = _torch()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    p = np.asarray(particles, dtype=np.float64)
    n, nx = p.shape
    ld = _round4(n)
    src = torch.zeros((nx, ld), dtype=torch.float64, device=dev)
    src[:, :n].copy_(torch.from_numpy(np.ascontiguousarray(p.T)))
    dst = torch.zeros_like(src)
    ix = torch.from_numpy(np.asarray(idx, dtype=np.int32)).to(dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    check(lib.smc_gather(C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), C.c_void_p(ix.data_ptr()),
                         n, ld, nx, _lib.SMC_F64, st))
    return dst[:, :n].cpu().numpy().T.copy()
