"""Multi-GPU layout of the batch: images are independent, so the index range is
cut into contiguous shards, one per rank (one process per GPU), each rank runs the whole
pipeline on its shard, and the only exchange is a host-side gather of the polygon lists in
image order.  No data-path collective exists or is needed."""
from __future__ import annotations


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous [first, first+count) of rank's shard; the first n_items % world ranks get one more."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_items, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def gather_polygon_scores(polygons, scores, dst=0, group=None):
    """polygons / scores: this rank's per-image lists (PolygonScores fields).  Returns the
    concatenation over ranks in rank (= image index) order on `dst`, None elsewhere.
    Works on any torch.distributed backend (the payload is a few KB per image)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(polygons), list(scores)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    payload = (list(polygons), list(scores))
    out = [None] * world if rank == dst else None
    dist.gather_object(payload, out, dst=dst, group=group)
    if rank != dst:
        return None
    all_p, all_s = [], []
    for p, s in out:
        all_p.extend(p)
        all_s.extend(s)
    return all_p, all_s


def gather_polygons(result, dst=0, group=None):
    """result: this rank's _ffi.Polygons (flat arrays).  Returns the _ffi.Polygons of the whole
    batch on `dst` (shards concatenated in rank = image index order), None elsewhere.  The
    payload is the five flat arrays, a few bytes per polygon point."""
    import torch.distributed as dist
    from ._ffi import Polygons
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return result
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    out = [None] * world if rank == dst else None
    dist.gather_object(result.arrays(), out, dst=dst, group=group)
    return Polygons.concat(out) if rank == dst else None


class Shards:
    """ocrb_shards: one ctx + detector + recognition net per device, driven by one host thread per device INSIDE
    libocrb (ocrb_detect_and_recognize_sharded, include/ocrb.h) — the single-process multi-GPU form of the batch
    contract text_detection/mod.rs:188-204."""

    def __init__(self, devices, det_weights, rec_weights=None, mode="bf16"):
        import ctypes as C

        from . import _ffi
        self._ffi = _ffi
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        nd, dn, dd, dl, keep_d = _ffi.weights_to_c(det_weights)
        if rec_weights is not None:
            nr, rn, rd, rl, keep_r = _ffi.weights_to_c(rec_weights)
        else:
            nr, rn, rd, rl, keep_r = 0, None, None, None, None
        self._h = _ffi.c_p()
        _ffi.check(_ffi.lib().ocrb_shards_create(devs, len(devices), nd, dn, dd, dl, _ffi.MODE_BF16 if mode == "bf16" else _ffi.MODE_FP32,
                                                 nr, rn, rd, rl, C.byref(self._h)))
        self.devices = list(devices)

    @property
    def launch_count(self):
        return int(self._ffi.lib().ocrb_shards_launch_count(self._h))

    def detect_and_recognize(self, images, adjust, glyphs=None, params=None):
        """images u8 [B,H,W] (host; pinned for full overlap), adjust f64 [B,2], glyphs u8 [n,784] or None
        -> (_ffi.Polygons of the whole batch in image order, glyph argmax int32 [n] or None)"""
        import ctypes as C

        import numpy as np
        _ffi = self._ffi
        B, H, W = images.shape
        adjust = np.ascontiguousarray(adjust, np.float64)
        n_gl = 0 if glyphs is None else len(glyphs)
        am = np.empty(n_gl, np.int32) if n_gl else None
        h = _ffi.c_p()
        _ffi.check(_ffi.lib().ocrb_detect_and_recognize_sharded(self._h, _ffi.ptr(images), _ffi.ptr(adjust), B, H, W, params,
                                                                _ffi.ptr(glyphs), n_gl, _ffi.ptr(am), C.byref(h)))
        return _ffi.Polygons(h), am

    def detect_and_read(self, images, adjust, glyphs_per_polygon=4, params=None):
        """ocrb_detect_and_read_sharded: polygons + glyph_classes [n_polygons, glyphs_per_polygon] of the whole batch"""
        import ctypes as C

        import numpy as np
        _ffi = self._ffi
        B, H, W = images.shape
        adjust = np.ascontiguousarray(adjust, np.float64)
        h = _ffi.c_p()
        _ffi.check(_ffi.lib().ocrb_detect_and_read_sharded(self._h, _ffi.ptr(images), _ffi.ptr(adjust), B, H, W, params,
                                                           int(glyphs_per_polygon), C.byref(h)))
        return _ffi.Polygons(h)

    def close(self):
        if self._h:
            self._ffi.lib().ocrb_shards_destroy(self._h)
            self._h = self._ffi.c_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShmGather:
    """Host-side gather of the per-rank polygon lists for the one-process-per-GPU layout, through
    POSIX shared memory on the box: rank r publishes the five flat arrays of its shard into its own segment
    (two slots, sequence-numbered), rank 0 collects the shards in rank (= image index) order.  No pickling, no
    collective, no GPU synchronisation: the cost is one memcpy of a few bytes per polygon point.

    key must be the same on every rank of the job and unique per job (e.g. MASTER_PORT)."""

    HDR = 8  # int64 words per slot header: seq, n_images, n_polys, n_points, + spare

    def __init__(self, rank, world, key, cap_bytes=16 << 20):
        import numpy as np
        self.np = np
        self.rank, self.world, self.cap = rank, world, cap_bytes
        self.paths = [f"/dev/shm/ocrb_gather_{key}_{r}.bin" for r in range(world)]
        words = 2 + 2 * (self.HDR + cap_bytes // 8)
        self.mine = np.memmap(self.paths[rank], dtype=np.int64, mode="w+", shape=(words,))
        self.mine[:] = 0
        self.mine[0] = -1  # ack: last step rank 0 has read from this segment
        self.mine.flush()
        self.all = None

    def _slot(self, seg, step):
        off = 2 + (step % 2) * (self.HDR + self.cap // 8)
        return seg[off:off + self.HDR], seg[off + self.HDR:off + self.HDR + self.cap // 8]

    def publish(self, result, step):
        """result: this rank's _ffi.Polygons; step: 0, 1, 2, ... (the same on all ranks)"""
        import time
        np = self.np
        io, po, xy, sc, st = result.arrays()
        gc = result.glyph_classes if result.glyph_classes is not None else np.zeros((0, 0), np.int32)
        while step >= 2 and self.mine[0] < step - 2:  # the reader still owns this slot
            time.sleep(20e-6)
        hdr, body = self._slot(self.mine, step)
        need = len(io) + len(po) + (xy.size + 1) // 2 + len(sc) + st.size + (gc.size + 1) // 2
        if need > len(body):
            raise RuntimeError(f"ShmGather: shard needs {need * 8} bytes, capacity {self.cap}")
        o = 0
        for a in (io, po):
            body[o:o + len(a)] = a
            o += len(a)
        body[o:o + (xy.size + 1) // 2].view(np.uint32)[:xy.size] = xy.reshape(-1)
        o += (xy.size + 1) // 2
        body[o:o + len(sc)].view(np.float64)[:] = sc
        o += len(sc)
        body[o:o + st.size] = st.reshape(-1)
        o += st.size
        body[o:o + (gc.size + 1) // 2].view(np.int32)[:gc.size] = gc.reshape(-1)
        hdr[1], hdr[2], hdr[3], hdr[4] = len(io) - 1, len(sc), len(xy), (gc.shape[1] if gc.size else 0)
        hdr[0] = step + 1  # published last: the slot is complete when the reader sees it

    def collect(self, step, timeout_s=60.0):
        """rank 0: the whole batch's _ffi.Polygons for `step` (blocks until every rank has published it)"""
        import os
        import time

        from ._ffi import Polygons
        np = self.np
        assert self.rank == 0
        t0 = time.time()
        if self.all is None:
            self.all = [self.mine]
            for r in range(1, self.world):
                while not os.path.exists(self.paths[r]) or os.path.getsize(self.paths[r]) < self.mine.nbytes:
                    if time.time() - t0 > timeout_s:
                        raise TimeoutError(f"ShmGather: rank {r} never created its segment")
                    time.sleep(1e-3)
                self.all.append(np.memmap(self.paths[r], dtype=np.int64, mode="r+", shape=self.mine.shape))
        parts = []
        for r, seg in enumerate(self.all):
            hdr, body = self._slot(seg, step)
            while hdr[0] != step + 1:
                if time.time() - t0 > timeout_s:
                    raise TimeoutError(f"ShmGather: rank {r} did not publish step {step}")
                time.sleep(20e-6)
            ni, npoly, npts, k = int(hdr[1]), int(hdr[2]), int(hdr[3]), int(hdr[4])
            o = 0
            io = np.array(body[o:o + ni + 1]); o += ni + 1
            po = np.array(body[o:o + npoly + 1]); o += npoly + 1
            xy = np.array(body[o:o + (2 * npts + 1) // 2].view(np.uint32)[:2 * npts]).reshape(-1, 2); o += (2 * npts + 1) // 2
            sc = np.array(body[o:o + npoly].view(np.float64)); o += npoly
            st = np.array(body[o:o + 5 * ni]).reshape(ni, 5); o += 5 * ni
            gc = np.array(body[o:o + (npoly * k + 1) // 2].view(np.int32)[:npoly * k]).reshape(npoly, k) if k else None
            parts.append((io, po, xy, sc, st, gc))
            seg[0] = step  # ack
        return Polygons.concat(parts)

    def close(self):
        import os
        try:
            del self.mine
            self.all = None
            os.unlink(self.paths[self.rank])
        except OSError:
            pass
