"""Host side of the engine: owns a C-ABI handle, moves buffers with torch (device memory, streams) and
calls the CUDA library.  PyTorch is plumbing only -- no torch op computes anything on the path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib


class MgicpError(RuntimeError):
    pass


@dataclass
class BatchResult:
    """Per-pair results of a batch; fields follow Open3D's RegistrationResult (AF:313, S2:198,218)."""
    transformation: np.ndarray        # [B,4,4]
    fitness: np.ndarray               # [B]
    inlier_rmse: np.ndarray           # [B]
    iterations: np.ndarray            # [B,S]
    num_correspondences: np.ndarray   # [B]
    stats: np.ndarray                 # [B,S,8]: M'_src, M'_tgt, iterations, K_last, fitness, rmse, sum K over passes, passes
    h2d_bytes: int = 0
    d2h_bytes: int = 0


@dataclass
class CloudFeatures:
    """Clouds with their normals and FPFH descriptors, resident in device memory (Engine.fpfh_clouds(..., resident=True))."""
    xyz: torch.Tensor            # flat coordinates as uploaded (float32 or float64, see code)
    off: np.ndarray              # [n_clouds + 1] point offsets (host)
    code: int                    # _lib.F32 / _lib.F64
    normals: torch.Tensor        # [total, 3] float64
    fpfh: torch.Tensor           # [total, 33] float64

    def host(self):
        """(list of [n,3] normals, list of [n,33] descriptors) on the host"""
        n_h, f_h, off = self.normals.cpu().numpy(), self.fpfh.cpu().numpy(), self.off
        return ([n_h[off[c]:off[c + 1]] for c in range(len(off) - 1)], [f_h[off[c]:off[c + 1]] for c in range(len(off) - 1)])


def _raise(L, h, rc, what):
    msg = L.mgicp_last_error(h)
    msg = msg.decode() if msg else ""
    text = f"{what}: {_lib.STATUS.get(rc, rc)}: {msg}"
    if rc in (1, 4):   # Open3D raises RuntimeError for voxel_size <= 0, max_correspondence_distance <= 0, voxel size too small
        raise RuntimeError(text)
    raise MgicpError(text)


class Engine:
    """One engine per (device, host thread), like the C handle it wraps."""

    def __init__(self, device: int | None = None):
        if not torch.cuda.is_available():
            raise MgicpError("mgicp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.L = _lib.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.tdev = torch.device("cuda", self.device)
        h = C.c_void_p()
        rc = self.L.mgicp_create(self.device, C.byref(h))
        if rc != 0:
            raise MgicpError(f"mgicp_create failed ({_lib.STATUS.get(rc, rc)})")
        self.h = h
        self._pinned = {}

    def close(self):
        if getattr(self, "h", None):
            self.L.mgicp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers -------------------------------------------------------------------------------
    def make_opts(self, *, sor_k=30, sor_std=1.0, normal_k=20, epsilon=1e-3, loss="l1", loss_k=1.0, rel_fitness=1e-6,
                  rel_rmse=1e-6, cell_factor=0.0, icp_cell_factor=0.0, ctas_per_pair=0, debug=False) -> _lib.Opts:
        if loss not in _lib.LOSS:
            raise ValueError(f"unknown loss {loss!r}")
        return _lib.Opts(int(sor_k), float(sor_std), int(normal_k), float(epsilon), _lib.LOSS[loss], float(loss_k),
                         float(rel_fitness), float(rel_rmse), float(cell_factor), float(icp_cell_factor), int(ctas_per_pair),
                         int(bool(debug)))

    def schedule_opts(self, opts: _lib.Opts, voxel_sizes, max_dists) -> _lib.Opts:
        """`opts` with the ICP grid's cell factor chosen for the schedule when the caller left it at 0 (mgicp_auto_icp_cell_factor:
        3.5 voxels for the script-2 schedule, up to 16 when the search radius is many voxels, as in ALL_FUNCTIONS.py:260-278)."""
        if opts.icp_cell_factor != 0.0:
            return opts
        vs = np.ascontiguousarray(voxel_sizes, np.float64)
        md = np.ascontiguousarray(max_dists, np.float64)
        md = np.ascontiguousarray(np.broadcast_to(md, (1, len(vs))) if md.ndim == 1 else md.reshape(-1, len(vs)))
        dp = C.POINTER(C.c_double)
        f = float(self.L.mgicp_auto_icp_cell_factor(len(vs), vs.ctypes.data_as(dp), md.shape[0], md.ctypes.data_as(dp)))
        out = _lib.Opts()
        C.memmove(C.byref(out), C.byref(opts), C.sizeof(_lib.Opts))
        out.icp_cell_factor = f
        return out

    def kernel_launches(self) -> int:
        return int(self.L.mgicp_kernel_launches(self.h))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.tdev).cuda_stream)

    @staticmethod
    def pack_clouds(clouds, dtype=None):
        """Concatenate N_i x 3 arrays -> (host array [sum N, 3], int64 offsets [C+1], dtype code)."""
        arrs = [np.asarray(getattr(c, "points", c)) for c in clouds]
        for a in arrs:
            if a.ndim != 2 or a.shape[1] != 3:
                raise ValueError("clouds must be N x 3")
        if dtype is None:
            dtype = np.float32 if all(a.dtype == np.float32 for a in arrs) else np.float64
        off = np.zeros(len(arrs) + 1, np.int64)
        off[1:] = np.cumsum([a.shape[0] for a in arrs])
        flat = np.empty((int(off[-1]), 3), dtype)
        for a, lo, hi in zip(arrs, off[:-1], off[1:]):
            flat[lo:hi] = a
        return flat, off, (_lib.F32 if dtype == np.float32 else _lib.F64)

    def upload(self, host: np.ndarray) -> torch.Tensor:
        """Pinned staging + async copy on the current stream.  The staging buffers are cached per (dtype, size); a buffer is
        only rewritten after the copy that last read it has completed (event per buffer)."""
        t = torch.from_numpy(np.ascontiguousarray(host))
        key = (t.dtype, t.numel())
        ent = self._pinned.get(key)
        if ent is None:
            if len(self._pinned) > 8:
                for _, ev in self._pinned.values():
                    ev.synchronize()
                self._pinned.clear()
            ent = (torch.empty(t.shape, dtype=t.dtype, pin_memory=True), torch.cuda.Event())
            self._pinned[key] = ent
        pin, ev = ent
        ev.synchronize()                      # a previous upload from this buffer may still be in flight
        pin = pin.view(t.shape)
        pin.copy_(t)
        dev = pin.to(self.tdev, non_blocking=True)
        ev.record(torch.cuda.current_stream(self.tdev))
        return dev

    # ---- stages --------------------------------------------------------------------------------
    def preprocess_device(self, xyz_dev: torch.Tensor, cloud_off: np.ndarray, voxel_sizes, opts: _lib.Opts):
        code = _lib.F32 if xyz_dev.dtype == torch.float32 else _lib.F64
        off = np.ascontiguousarray(cloud_off, np.int64)
        vs = np.ascontiguousarray(voxel_sizes, np.float64)
        rc = self.L.mgicp_preprocess(self.h, self._stream(), len(off) - 1, C.c_void_p(xyz_dev.data_ptr()),
                                     off.ctypes.data_as(C.POINTER(C.c_int64)), code, len(vs),
                                     vs.ctypes.data_as(C.POINTER(C.c_double)), C.byref(opts))
        if rc != 0:
            _raise(self.L, self.h, rc, "mgicp_preprocess")
        self._n_scales = len(vs)

    def register_device(self, pair_src, pair_tgt, max_dists, max_iters, T_init_dev: torch.Tensor, opts: _lib.Opts):
        """ICP loops for all pairs; returns device tensors (T, fitness, rmse, iters, ncorr, stats). Asynchronous."""
        ps = np.ascontiguousarray(pair_src, np.int32)
        pt = np.ascontiguousarray(pair_tgt, np.int32)
        B, S = len(ps), self._n_scales
        md = np.ascontiguousarray(max_dists, np.float64).reshape(B, S)
        mi = np.ascontiguousarray(max_iters, np.int32).reshape(S)
        T = torch.empty((B, 4, 4), dtype=torch.float64, device=self.tdev)
        fit = torch.empty((B,), dtype=torch.float64, device=self.tdev)
        rm = torch.empty((B,), dtype=torch.float64, device=self.tdev)
        it = torch.zeros((B, S), dtype=torch.int32, device=self.tdev)
        nc = torch.zeros((B,), dtype=torch.int32, device=self.tdev)
        st = torch.zeros((B, S, 8), dtype=torch.float64, device=self.tdev)
        i32p, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
        rc = self.L.mgicp_register_batch(self.h, self._stream(), B, ps.ctypes.data_as(i32p), pt.ctypes.data_as(i32p),
                                         md.ctypes.data_as(dp), mi.ctypes.data_as(i32p), C.byref(opts),
                                         C.c_void_p(T_init_dev.data_ptr()), C.c_void_p(T.data_ptr()), C.c_void_p(fit.data_ptr()),
                                         C.c_void_p(rm.data_ptr()), C.c_void_p(it.data_ptr()), C.c_void_p(nc.data_ptr()),
                                         C.c_void_p(st.data_ptr()))
        if rc != 0:
            _raise(self.L, self.h, rc, "mgicp_register_batch")
        return T, fit, rm, it, nc, st

    def check(self):
        rc = self.L.mgicp_check(self.h)
        if rc != 0:
            _raise(self.L, self.h, rc, "mgicp_check")

    def job_errors_device(self) -> torch.Tensor:
        """stream-ordered mgicp_check: an int32[1] device tensor with the first error flag of the last preprocess (0 = none)"""
        err = torch.zeros((1,), dtype=torch.int32, device=self.tdev)
        rc = self.L.mgicp_job_errors(self.h, self._stream(), C.c_void_p(err.data_ptr()))
        if rc != 0:
            _raise(self.L, self.h, rc, "mgicp_job_errors")
        return err

    def raise_job_error(self, code: int):
        if code:
            text = ("extent / voxel_size exceeds 2^21 cells per axis" if code == 4 else "internal hash table overflow")
            if code == 4:
                raise RuntimeError(f"mgicp: RANGE: {text}")
            raise MgicpError(f"mgicp: {_lib.STATUS.get(code, code)}: {text}")

    def correspondences(self, pair: int, n_cap: int) -> np.ndarray:
        """[K, 2] int32 correspondence_set of pair `pair` of the last register call (indices into the last scale's final clouds)"""
        buf = np.empty((max(int(n_cap), 1), 2), np.int32)
        cnt = C.c_int64(0)
        rc = self.L.mgicp_get_correspondences(self.h, int(pair), buf.ctypes.data_as(C.c_void_p), buf.shape[0], C.byref(cnt))
        if rc != 0:
            _raise(self.L, self.h, rc, "mgicp_get_correspondences")
        return buf[: cnt.value].copy()

    def set_timing(self, on: bool = True):
        self.L.mgicp_set_timing(self.h, int(bool(on)))

    def get_timing(self) -> dict:
        """per-stage milliseconds of the last preprocess (+ register) on this engine, see mgicp_get_timing in include/mgicp.h"""
        out = (C.c_double * 16)()
        rc = self.L.mgicp_get_timing(self.h, out)
        if rc != 0:
            _raise(self.L, self.h, rc, "mgicp_get_timing")
        v = list(out)
        return {"downsample_ms": v[0], "knn_grid_ms": v[1], "sor_ms": v[2], "normals_ms": v[3], "icp_grid_ms": v[4], "icp_ms": v[5],
                "scale_ms": v[8:16]}

    def evaluate(self, scale, pair_src, pair_tgt, max_dists, T, opts):
        """One correspondence pass (evaluate_registration, AF:809-822) + the GICP normal equations at pose T."""
        ps = np.ascontiguousarray(pair_src, np.int32)
        pt = np.ascontiguousarray(pair_tgt, np.int32)
        md = np.ascontiguousarray(max_dists, np.float64).reshape(len(ps))
        Td = self.upload(np.ascontiguousarray(T, np.float64).reshape(len(ps), 16))
        out = torch.zeros((len(ps), 32), dtype=torch.float64, device=self.tdev)
        i32p, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
        rc = self.L.mgicp_evaluate_batch(self.h, self._stream(), int(scale), len(ps), ps.ctypes.data_as(i32p),
                                         pt.ctypes.data_as(i32p), md.ctypes.data_as(dp), C.byref(opts), C.c_void_p(Td.data_ptr()),
                                         C.c_void_p(out.data_ptr()))
        if rc != 0:
            _raise(self.L, self.h, rc, "mgicp_evaluate_batch")
        o = out.cpu().numpy()
        return dict(fitness=o[:, 0], rmse=o[:, 1], K=o[:, 2], sum_d2=o[:, 3], sums=o[:, 4:31])

    def evaluate_clouds(self, clouds, pairs, max_dists, T, want_corr=False):
        """evaluate_registration + get_information_matrix_from_point_clouds on the clouds as given, for a batch of
        (source_index, target_index) pairs (AF:809-822, AF:327-331).  Returns dict(fitness, rmse, K, sum_d2, information
        [B,6,6] (+ corr: list of int32 arrays))."""
        flat, off, code = self.pack_clouds(clouds)
        B = len(pairs)
        ps = np.ascontiguousarray([p[0] for p in pairs], np.int32)
        pt = np.ascontiguousarray([p[1] for p in pairs], np.int32)
        md = np.ascontiguousarray(np.broadcast_to(np.asarray(max_dists, np.float64), (B,)))
        Th = np.ascontiguousarray(T, np.float64).reshape(B, 16)
        xyz = self.upload(flat)
        out = torch.zeros((B, 32), dtype=torch.float64, device=self.tdev)
        ns = [int(off[s + 1] - off[s]) for s in ps]
        corr = torch.empty((max(1, sum(ns)),), dtype=torch.int32, device=self.tdev) if want_corr else None
        i32p, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
        rc = self.L.mgicp_evaluate_clouds(self.h, self._stream(), len(off) - 1, C.c_void_p(xyz.data_ptr()),
                                          off.ctypes.data_as(C.POINTER(C.c_int64)), code, B, ps.ctypes.data_as(i32p),
                                          pt.ctypes.data_as(i32p), md.ctypes.data_as(dp), Th.ctypes.data_as(dp),
                                          C.c_void_p(out.data_ptr()), C.c_void_p(corr.data_ptr()) if want_corr else None)
        if rc != 0:
            _raise(self.L, self.h, rc, "mgicp_evaluate_clouds")
        o = out.cpu().numpy()
        self.check()
        info = np.zeros((B, 6, 6))
        iu = np.triu_indices(6)
        info[:, iu[0], iu[1]] = o[:, 4:25]
        info[:, iu[1], iu[0]] = o[:, 4:25]
        res = dict(fitness=o[:, 0], rmse=o[:, 1], K=o[:, 2], sum_d2=o[:, 3], information=info)
        if want_corr:
            c = corr.cpu().numpy()
            cuts = np.cumsum([0] + ns)
            res["corr"] = [c[cuts[b]:cuts[b + 1]] for b in range(B)]
        return res

    def fpfh_clouds(self, clouds, radius_normals, max_nn_normals, radius_fpfh, max_nn_fpfh, *, resident=False):
        """estimate_normals(Hybrid(radius_normals, max_nn_normals)) + compute_fpfh_feature(Hybrid(radius_fpfh, max_nn_fpfh)) for
        every cloud as given (AF:181-187).  Returns (list of [n,3] normals, list of [n,33] descriptors); with `resident=True`
        a CloudFeatures instead: clouds, normals and descriptors stay in device memory for `fgr_pairs` (33 doubles per point
        do not travel to the host and back).
        CUDA path of the FGR front end (csrc/mgicp_fgr.cuh): normals and descriptors bit-identical to the oracle."""
        flat, off, code = self.pack_clouds(clouds)
        xyz = self.upload(flat)
        total = int(off[-1])
        nrm = torch.empty((max(1, total), 3), dtype=torch.float64, device=self.tdev)
        fp = torch.empty((max(1, total), 33), dtype=torch.float64, device=self.tdev)
        rc = self.L.mgicp_fpfh_clouds(self.h, self._stream(), len(off) - 1, C.c_void_p(xyz.data_ptr()),
                                      off.ctypes.data_as(C.POINTER(C.c_int64)), code, float(radius_normals), int(max_nn_normals),
                                      float(radius_fpfh), int(max_nn_fpfh), C.c_void_p(nrm.data_ptr()), C.c_void_p(fp.data_ptr()))
        if rc != 0:
            _raise(self.L, self.h, rc, "mgicp_fpfh_clouds")
        if resident:
            return CloudFeatures(xyz, off, code, nrm, fp)
        n_h, f_h = nrm.cpu().numpy(), fp.cpu().numpy()
        self.check()
        return ([n_h[off[c]:off[c + 1]] for c in range(len(off) - 1)], [f_h[off[c]:off[c + 1]] for c in range(len(off) - 1)])

    def fgr_pairs(self, clouds, features, pairs, *, division_factor=1.4, use_absolute_scale=False, decrease_mu=False,
                  maximum_correspondence_distance=0.025, iteration_number=64, tuple_scale=0.95, maximum_tuple_count=1000,
                  seeds=None):
        """registration_fgr_based_on_feature_matching (AF:196-201) for a batch of (source_index, target_index) pairs over
        clouds with their [n, 33] descriptors; keyword defaults are Open3D's FastGlobalRegistrationOption.  `features` is a
        list of host arrays, or the CloudFeatures of `fpfh_clouds(..., resident=True)` (then `clouds` is not read again).
        Returns (T [B,4,4] source -> target, number of correspondences optimised [B]).
        Matching on the tensor cores with an exact fp64 re-check (csrc/mgicp_fgr_tc.cuh), then one block per pair."""
        if isinstance(features, CloudFeatures):
            xyz, off, code, fdev = features.xyz, features.off, features.code, features.fpfh
        else:
            flat, off, code = self.pack_clouds(clouds)
            feat = np.ascontiguousarray(np.concatenate([np.asarray(f, np.float64).reshape(-1, 33) for f in features]), np.float64)
            if feat.shape[0] != int(off[-1]):
                raise ValueError("one 33-bin descriptor per point is required")
            xyz, fdev = self.upload(flat), self.upload(feat)
        B = len(pairs)
        ps = np.ascontiguousarray([p[0] for p in pairs], np.int32)
        pt = np.ascontiguousarray([p[1] for p in pairs], np.int32)
        sd = np.ascontiguousarray(np.arange(B) if seeds is None else seeds, np.uint64)
        caps = None if np.isscalar(maximum_tuple_count) else np.ascontiguousarray(maximum_tuple_count, np.int32)   # per pair
        if caps is not None and caps.shape != (B,):
            raise ValueError("maximum_tuple_count: one value, or one per pair")
        o = _lib.FgrOpts(division_factor, int(use_absolute_scale), int(decrease_mu), maximum_correspondence_distance, iteration_number,
                         tuple_scale, int(maximum_tuple_count) if caps is None else int(caps.max(initial=0)))
        T = torch.zeros((B, 16), dtype=torch.float64, device=self.tdev)
        nc = torch.zeros((B,), dtype=torch.int32, device=self.tdev)
        i32p = C.POINTER(C.c_int32)
        rc = self.L.mgicp_fgr_pairs(self.h, self._stream(), len(off) - 1, C.c_void_p(xyz.data_ptr()), off.ctypes.data_as(C.POINTER(C.c_int64)),
                                    code, C.c_void_p(fdev.data_ptr()), B, ps.ctypes.data_as(i32p), pt.ctypes.data_as(i32p), C.byref(o),
                                    None if caps is None else caps.ctypes.data_as(i32p), sd.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_void_p(T.data_ptr()), C.c_void_p(nc.data_ptr()))
        if rc != 0:
            _raise(self.L, self.h, rc, "mgicp_fgr_pairs")
        return T.cpu().numpy().reshape(B, 4, 4), nc.cpu().numpy()

    def get_stage(self, cloud: int, scale: int, what: int, n_cap: int, k: int = 1):
        if what in (_lib.STAGE_KNN_SOR, _lib.STAGE_KNN_NORMAL):
            buf = np.empty((n_cap, k), np.int32)
        elif what == _lib.STAGE_SOR_KEEP:
            buf = np.empty((n_cap,), np.uint8)
        elif what == _lib.STAGE_SOR_AVG:
            buf = np.empty((n_cap,), np.float64)
        elif what == _lib.STAGE_BOUNDS:
            buf = np.empty((6,), np.float64)
        else:
            buf = np.empty((n_cap, 3), np.float64)
        cnt = C.c_int64(0)
        rc = self.L.mgicp_get_stage(self.h, cloud, scale, what, buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(cnt))
        if rc != 0:
            _raise(self.L, self.h, rc, "mgicp_get_stage")
        return buf if what == _lib.STAGE_BOUNDS else buf[: cnt.value]

    def cloud_bounds(self, clouds):
        flat, off, code = self.pack_clouds(clouds)
        xyz = self.upload(flat)
        out = torch.empty((len(off) - 1, 6), dtype=torch.float64, device=self.tdev)
        rc = self.L.mgicp_cloud_bounds(self.h, self._stream(), len(off) - 1, C.c_void_p(xyz.data_ptr()),
                                       off.ctypes.data_as(C.POINTER(C.c_int64)), code, C.c_void_p(out.data_ptr()))
        if rc != 0:
            _raise(self.L, self.h, rc, "mgicp_cloud_bounds")
        return out.cpu().numpy()

    # ---- the whole path, host buffers in, host results out ------------------------------------------
    def run(self, clouds, pairs, voxel_sizes, max_dists, max_iters, T_init, opts: _lib.Opts | None = None) -> BatchResult:
        """clouds: list of N_i x 3 arrays; pairs: list of (source_index, target_index);
        max_dists: [S] or [B,S]; max_iters: int or [S]; T_init: [B,4,4]."""
        opts = opts or self.make_opts()
        flat, off, _ = self.pack_clouds(clouds)
        S, B = len(voxel_sizes), len(pairs)
        md = np.asarray(max_dists, np.float64)
        md = np.broadcast_to(md, (B, S)) if md.ndim == 1 else md.reshape(B, S)
        mi = np.full(S, int(max_iters), np.int32) if np.isscalar(max_iters) else np.asarray(max_iters, np.int32)
        T0 = np.ascontiguousarray(T_init, np.float64).reshape(B, 4, 4)
        xyz = self.upload(flat)
        T0d = self.upload(T0)
        opts = self.schedule_opts(opts, voxel_sizes, md)
        self.preprocess_device(xyz, off, voxel_sizes, opts)
        ps = [p[0] for p in pairs]
        pt = [p[1] for p in pairs]
        T, fit, rm, it, nc, st = self.register_device(ps, pt, md, mi, T0d, opts)
        Th, fh, rh, ih, nh, sh = (t.cpu().numpy() for t in (T, fit, rm, it, nc, st))
        self.check()
        h2d = flat.nbytes + T0.nbytes
        d2h = Th.nbytes + fh.nbytes + rh.nbytes + ih.nbytes + nh.nbytes + sh.nbytes
        return BatchResult(Th, fh, rh, ih, nh, sh, h2d, d2h)
