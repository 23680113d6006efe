"""Streaming form of the batched refinement: host clouds in, host poses out, batch after batch.

The reference's loop (2_MGICP_refinement_in_NCLT_dataset.py:187-218) hands pageable numpy clouds to
``Multiscale_GICP`` one pair at a time.  A caller who has many pairs gives them to ``BatchStream`` in
batches (lists of clouds + the pairs over them + the initial poses) and gets one ``BatchResult`` per
batch, in order.  Everything a real caller pays is inside: packing the clouds into pinned staging
memory (a worker thread, one copy, straight from the caller's arrays), the host-to-device copy on a
copy stream, preprocessing + the ICP launch on a work stream, and the device-to-host copy of the results
on a third stream.

Overlap (measured with bench.py's `e2e`): consecutive batches alternate between two engines (two
workspaces, two work streams), so the preprocessing kernels of batch k+1 fill the SMs that the
persistent ICP kernel of batch k leaves idle towards its end; the upload of batch k+1 is ordered after
the preprocessing kernels of batch k (it would slow those bandwidth- and atomics-heavy kernels down)
and runs under batch k's latency-bound ICP kernel; the host blocks only on the results of batch k-1,
after it has enqueued batch k.  PyTorch provides pinned memory, streams and events: plumbing only.
"""
from __future__ import annotations

import collections
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import _lib
from .engine import BatchResult, Engine


class _Slot:
    """one of the two staging areas: pinned host clouds / poses, their device copies, the pinned result block"""

    def __init__(self):
        self.pin_xyz = None      # pinned uint8 storage, viewed as float32 / float64 [n, 3]
        self.pin_T0 = None
        self.pin_res = None
        self.dev_xyz = None
        self.dev_T0 = None
        self.ev_up = torch.cuda.Event()       # H2D of this slot finished
        self.ev_free = torch.cuda.Event()     # the compute that read this slot's device buffers finished
        self.ev_step = torch.cuda.Event()     # the result block of this slot is complete on the device
        self.ev_out = torch.cuda.Event()      # ... and has landed in pinned host memory


def _grow_pinned(buf, nbytes):
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty((int(nbytes * 1.25) + 256,), dtype=torch.uint8, pin_memory=True)
    return buf


class BatchStream:
    """Multiscale GICP over a stream of batches.

        bs = BatchStream(voxel_sizes, max_dists, max_iters, device=0, loss="l1")
        for res in bs.run(batches):        # batches: iterable of (clouds, pairs, T_init[, max_dists_of_the_batch])
            ...                            # res: engine.BatchResult (numpy arrays), one per batch, in order

    clouds: list of N_i x 3 float32 / float64 arrays (or objects with .points); pairs: list of (source_index, target_index)
    into that list; T_init: [B, 4, 4].  Inputs are never modified (AF:289-290)."""

    def __init__(self, voxel_sizes, max_dists, max_iters, *, device: int | None = None, engines=2,
                 engine: Engine | None = None, opts: _lib.Opts | None = None, post=None, pack_threads: int = 4, **opt_kw):
        """engines: how many engines (workspaces) alternate, or a list of existing Engine objects to use;
        post: optional callable applied to the [B, cols] float64 device result block on the work stream before it is copied
        to the host (multi-GPU callers all-gather the blocks of all ranks there: one collective per batch, nothing else)."""
        if isinstance(engines, (list, tuple)):
            self.engs = list(engines)
            first = self.engs[0]
        else:
            first = engine or Engine(device)
            self.engs = [first] + [Engine(first.device) for _ in range(max(1, engines) - 1)]
        self.post = post
        self.tdev = first.tdev
        self.voxels = [float(v) for v in voxel_sizes]
        self.S = len(self.voxels)
        self.max_dists = np.asarray(max_dists, np.float64)
        self.opts = first.schedule_opts(opts or first.make_opts(**opt_kw), self.voxels, self.max_dists)
        self.max_iters = (np.full(self.S, int(max_iters), np.int32) if np.isscalar(max_iters)
                          else np.ascontiguousarray(max_iters, np.int32).reshape(self.S))
        self.work = [torch.cuda.Stream(device=self.tdev) for _ in self.engs]
        self.copy = torch.cuda.Stream(device=self.tdev)
        self.d2h = torch.cuda.Stream(device=self.tdev)
        self.slots = [_Slot(), _Slot()]
        self.pool = ThreadPoolExecutor(max_workers=1)                 # one batch at a time is being packed ...
        self.copiers = ThreadPoolExecutor(max_workers=pack_threads)   # ... by several threads (numpy copies release the GIL)
        self.pack_threads = pack_threads
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def close(self):
        self.pool.shutdown(wait=True)
        self.copiers.shutdown(wait=True)

    def kernel_launches(self) -> int:
        return sum(e.kernel_launches() for e in self.engs)

    # ---- host side: pack one batch into a slot's pinned memory (worker thread) ------------------------------------
    def _pack(self, slot: _Slot, batch):
        clouds, pairs, T_init = batch[0], batch[1], batch[2]
        arrs = [np.asarray(getattr(c, "points", c)) for c in clouds]
        for a in arrs:
            if a.ndim != 2 or a.shape[1] != 3:
                raise ValueError("clouds must be N x 3")
        dtype = np.float32 if all(a.dtype == np.float32 for a in arrs) else np.float64
        off = np.zeros(len(arrs) + 1, np.int64)
        off[1:] = np.cumsum([a.shape[0] for a in arrs])
        n = int(off[-1])
        nbytes = n * 3 * np.dtype(dtype).itemsize
        slot.ev_up.synchronize()                       # the last upload from this pinned area has been read
        slot.pin_xyz = _grow_pinned(slot.pin_xyz, max(nbytes, 16))
        flat = slot.pin_xyz[:nbytes].numpy().view(dtype).reshape(n, 3)
        def copy_range(c0, c1):                        # the one host copy: caller's arrays -> pinned staging
            for c in range(c0, c1):
                flat[off[c]:off[c + 1]] = arrs[c]
        # split the clouds into pack_threads groups of about equal size (a single thread moves ~4 GB/s into pinned memory: a
        # 350 MB batch would take longer to pack than the GPU needs to process it)
        nt = max(1, min(self.pack_threads, len(arrs)))
        cuts = np.searchsorted(off, np.linspace(0, n, nt + 1)[1:-1]).tolist() if nt > 1 else []
        bounds = [0] + cuts + [len(arrs)]
        futs = [self.copiers.submit(copy_range, bounds[g], bounds[g + 1]) for g in range(nt) if bounds[g + 1] > bounds[g]]
        for f in futs:
            f.result()
        B = len(pairs)
        T0 = np.ascontiguousarray(T_init, np.float64).reshape(B, 16)
        slot.pin_T0 = _grow_pinned(slot.pin_T0, T0.nbytes)
        slot.pin_T0[:T0.nbytes].numpy().view(np.float64).reshape(B, 16)[:] = T0
        md = np.asarray(batch[3], np.float64) if len(batch) > 3 and batch[3] is not None else self.max_dists
        md = np.broadcast_to(md, (B, self.S)).copy() if md.ndim == 1 else np.ascontiguousarray(md).reshape(B, self.S)
        return dict(off=off, n=n, nbytes=nbytes, dtype=dtype, B=B, md=md,
                    ps=np.ascontiguousarray([p[0] for p in pairs], np.int32), pt=np.ascontiguousarray([p[1] for p in pairs], np.int32))

    # ---- device side ---------------------------------------------------------------------------------------------------
    def _upload(self, slot: _Slot, meta, after: torch.cuda.Event | None):
        tdt = torch.float32 if meta["dtype"] == np.float32 else torch.float64
        with torch.cuda.stream(self.copy):
            if after is not None:
                self.copy.wait_event(after)
            self.copy.wait_event(slot.ev_free)         # the batch that last used this slot's device buffers is done
            if slot.dev_xyz is None or slot.dev_xyz.numel() < meta["nbytes"]:
                slot.dev_xyz = torch.empty((int(meta["nbytes"] * 1.25) + 256,), dtype=torch.uint8, device=self.tdev)
            if slot.dev_T0 is None or slot.dev_T0.numel() < meta["B"] * 128:
                slot.dev_T0 = torch.empty((meta["B"] * 128 + 256,), dtype=torch.uint8, device=self.tdev)
            slot.dev_xyz[:meta["nbytes"]].copy_(slot.pin_xyz[:meta["nbytes"]], non_blocking=True)
            slot.dev_T0[:meta["B"] * 128].copy_(slot.pin_T0[:meta["B"] * 128], non_blocking=True)
            slot.ev_up.record(self.copy)
        meta["xyz_dev"] = slot.dev_xyz[:meta["nbytes"]].view(tdt).view(meta["n"], 3)
        meta["T0_dev"] = slot.dev_T0[:meta["B"] * 128].view(torch.float64).view(meta["B"], 16)
        self.h2d_bytes += meta["nbytes"] + meta["B"] * 128

    def _fetch(self, slot: _Slot, meta) -> BatchResult:
        B, S = meta["B"], self.S
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(slot.ev_step)
            res = meta["res_dev"]
            slot.pin_res = _grow_pinned(slot.pin_res, res.numel() * 8)
            host = slot.pin_res[:res.numel() * 8].view(torch.float64).view(res.shape)
            host.copy_(res, non_blocking=True)
            slot.ev_out.record(self.d2h)
        slot.ev_out.synchronize()
        a = host.numpy().copy()
        self.d2h_bytes += a.nbytes
        B = a.shape[0]                                  # more rows than pairs of this batch if `post` gathered other ranks' blocks
        T = a[:, :16].reshape(B, 4, 4)
        it = np.rint(a[:, 18:18 + S]).astype(np.int32)
        nc = np.rint(a[:, 18 + S]).astype(np.int32)
        st = a[:, 19 + S:19 + 9 * S].reshape(B, S, 8)
        meta["engine"].raise_job_error(int(a[:, 19 + 9 * S].max()) if B else 0)
        return BatchResult(T, a[:, 16].copy(), a[:, 17].copy(), it, nc, st, meta["nbytes"] + meta["B"] * 128, a.nbytes)

    def run(self, batches):
        """generator: one BatchResult per batch, in order"""
        it = iter(batches)
        queue = collections.deque()                    # (future of the packed meta, slot index)
        n_in = 0

        def submit_next():
            nonlocal n_in
            try:
                b = next(it)
            except StopIteration:
                return False
            queue.append(self.pool.submit(self._pack, self.slots[n_in & 1], b))
            n_in += 1
            return True

        cur = torch.cuda.current_stream(self.tdev)
        for s in self.slots:
            s.ev_free.record(cur)
        if not submit_next():
            return
        meta = queue.popleft().result()
        self._upload(self.slots[0], meta, None)
        submit_next()
        pending = None
        k = 0
        ev_pre = torch.cuda.Event()
        while meta is not None:
            slot = self.slots[k & 1]
            e = k % len(self.engs)
            eng, ws = self.engs[e], self.work[e]
            ws.wait_event(slot.ev_up)
            with torch.cuda.stream(ws):
                eng.preprocess_device(meta["xyz_dev"], meta["off"], self.voxels, self.opts)
                ev_pre.record(ws)
                T, fit, rm, its, nc, st = eng.register_device(meta["ps"], meta["pt"], meta["md"], self.max_iters, meta["T0_dev"], self.opts)
                B = meta["B"]
                err = eng.job_errors_device()          # comes home with the results: no device-wide synchronisation
                meta["res_dev"] = torch.cat([T.reshape(B, 16), fit[:, None], rm[:, None], its.to(torch.float64),
                                             nc.to(torch.float64)[:, None], st.reshape(B, -1),
                                             err.to(torch.float64).expand(B)[:, None]], dim=1)
                if self.post is not None:
                    meta["res_dev"] = self.post(meta["res_dev"])
                slot.ev_step.record(ws)
                slot.ev_free.record(ws)
            meta["engine"] = eng
            # the next batch: packed while the previous one computed; its upload runs under this batch's ICP kernel
            nxt = None
            if queue:
                nxt = queue.popleft().result()
                self._upload(self.slots[(k + 1) & 1], nxt, ev_pre)
            if pending is not None:
                yield self._finish(*pending)           # blocks on batch k-1 only, with batch k already enqueued
            # the slot of batch k-1 is free for packing once its results are home (its upload completed long ago)
            submit_next()
            pending = (slot, meta)
            meta = nxt
            k += 1
        yield self._finish(*pending)

    def _finish(self, slot, meta):
        return self._fetch(slot, meta)

    def run_one(self, clouds, pairs, T_init, max_dists=None) -> BatchResult:
        return next(iter(self.run([(clouds, pairs, T_init, max_dists)])))
