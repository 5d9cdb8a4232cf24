"""Multi-GPU execution of the RRDB path: one process per GPU, no collective on the data path.

The path shards trivially: tiles of one frame, and frames, are independent.
Two modes:

* image sharding (``frames_for_rank``): rank r upscales frames r, r+G, ... on its own GPU --
  throughput mode (bench.py's weak-scaling leg);
* tile sharding (``TileShardedUpscaler``): the ranks split the row-major tile list of ONE frame into
  contiguous ranges; every rank's last conv stores its finished tiles straight into the frame owner's
  tile buffer through a CUDA-IPC peer mapping (NVLink/NVSwitch P2P stores), then the owner runs the
  single gather-blend kernel in fixed tile order, so the result is bit-identical for any number of
  ranks (bench.py's strong-scaling leg).  The reference analogue is the serial tile loop of
  run.py:187-202.

Tile sharding is a pipeline with no host synchronisation per frame.  Ownership rotates (frame f ->
rank f mod G); three streams per rank -- upload, compute, blend -- are ordered against the other ranks
by monotonically increasing device counters that every rank keeps in a small table its peers map
through CUDA IPC (csrc/sync_ops.cu: a signal is a release store after the stream's earlier work, a
wait holds a stream until the counters reach a value).  With ``v = f + 1`` for frame f:

    owner, upload stream : wait TILES_DONE[all] >= f-1      (LR slot f%2 is free: frame f-2 was read)
                           H2D pinned frame -> own LR slot, copy to every peer's LR slot (NVLink)
                           signal LR_READY[owner] = v  in every rank's table
    all,  compute stream : wait LR_READY[owner] >= v,  wait BLEND_DONE[owner] >= f-G+1 (tile buffer free)
                           tiles [t0, t1) of the frame -> owner's tile buffer (peer stores from the last conv)
                           signal TILES_DONE[rank] = v  in every rank's table
    owner, blend stream  : wait TILES_DONE[all] >= v;  blend -> uint8;  signal BLEND_DONE[owner] = v;
                           D2H into pinned host memory

so the owner's blend and D2H of frame f overlap everybody's compute of frame f+1, and the upload of
frame f+1 overlaps the compute of frame f.  torch.distributed is used for the handle exchange only.
"""
import ctypes

import numpy as np

LR_READY, TILES_DONE, BLEND_DONE = 0, 1, 2
MAX_RANKS = 16          # flag table: [3 kinds][MAX_RANKS] uint32 + one error word


def partition(n_items, world):
    """Contiguous ranges of ceil(n/world) items: [(begin, end)] * world."""
    per = -(-n_items // world)
    return [(min(r * per, n_items), min((r + 1) * per, n_items)) for r in range(world)]


def frame_owner(frame_index, world):
    return frame_index % world


def frames_for_rank(n_frames, rank, world):
    return list(range(rank, n_frames, world))


def flag_index(kind, src):
    return kind * MAX_RANKS + src


class NativeTileBackend:
    """Device-side operations of the tile-sharded mode on top of the C-ABI (include/innfer_b200.h).

    Owns, on this rank's GPU: two LR frame slots, one tile buffer (dedicated, never resized -- peers hold
    IPC mappings of it), the flag table, a uint8 output frame, two pinned host result slots and the three
    streams.  All methods only enqueue work; ``fetch`` is the one that blocks."""

    def __init__(self, engine, H, W, patch=200, step=0.5, wait_timeout_ms=20000):
        import torch
        from . import _native as N
        self.N, self.lib, self.eng, self.torch = N, N.load(), engine, torch
        self.H, self.W, self.patch, self.step = H, W, patch, float(step)
        self.scale = engine.cfg["scale"]
        self.timeout_ms = int(wait_timeout_ms)
        self.dev = torch.device("cuda", engine.index)
        self._opened = []
        total, per_tile, n = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_int()
        N.check(self.lib.innfer_rrdb_tile_bytes(engine._h, H, W, patch, self.step, ctypes.byref(total), ctypes.byref(per_tile),
                                                ctypes.byref(n)))
        self.ntiles, self.tile_bytes = n.value, per_tile.value
        self.lr_bytes = H * W * 3
        self.lr_ptr = self._alloc(2 * self.lr_bytes)
        self.tiles_ptr = self._alloc(total.value)
        self.flags_bytes = (3 * MAX_RANKS + 1) * 4
        self.flags_ptr = self._alloc(self.flags_bytes)
        N.check(self.lib.innfer_device_memset(ctypes.c_void_p(self.flags_ptr), 0, self.flags_bytes))
        with torch.cuda.device(self.dev):
            self.s_up, self.s_cmp, self.s_bl = (torch.cuda.Stream(self.dev) for _ in range(3))
            self.d_out = torch.empty((self.scale * H, self.scale * W, 3), dtype=torch.uint8, device=self.dev)
        self.h_out = [torch.empty((self.scale * H, self.scale * W, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.h_in = [torch.empty((H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._stage_ev = [None, None]   # completion of the last H2D out of each staging slot
        # Launch both flag kernels once while nothing blocks: with CUDA's lazy module loading the FIRST launch of a
        # kernel may have to wait for running kernels, and a spinning wait kernel is exactly what must not be waited
        # for (measured: a signal first launched behind a running wait arrived only after the wait's timeout).
        scratch = (ctypes.c_void_p * 1)(self.flags_ptr + 4 * 3 * MAX_RANKS)
        with torch.cuda.device(self.dev):
            N.check(self.lib.innfer_stream_signal(scratch, 1, 0, ctypes.c_void_p(self.s_up.cuda_stream)))
            N.check(self.lib.innfer_stream_wait(scratch, 1, 0, None, 1000, ctypes.c_void_p(self.s_up.cuda_stream)))
            torch.cuda.synchronize(self.dev)
        self._done = {}
        self._nown = 0

    def _alloc(self, nbytes):
        p = ctypes.c_void_p()
        self.N.check(self.lib.innfer_device_alloc(self.eng.index, nbytes, ctypes.byref(p)))
        return p.value

    # -- handle exchange --------------------------------------------------------------------------
    def export_handles(self):
        out = []
        for p in (self.lr_ptr, self.tiles_ptr, self.flags_ptr):
            buf = (ctypes.c_uint8 * 64)()
            self.N.check(self.lib.innfer_ipc_export(ctypes.c_void_p(p), buf))
            out.append(bytes(buf))
        return out

    def open_handles(self, handles):
        ptrs = []
        with self.torch.cuda.device(self.dev):
            for hd in handles:
                buf = (ctypes.c_uint8 * 64).from_buffer_copy(hd)
                p = ctypes.c_void_p()
                self.N.check(self.lib.innfer_ipc_open(buf, ctypes.byref(p)))
                self._opened.append(p.value)
                ptrs.append(p.value)
        return ptrs

    def local_ptrs(self):
        return [self.lr_ptr, self.tiles_ptr, self.flags_ptr]

    # -- stream-ordered operations ----------------------------------------------------------------
    def _stream(self, which):
        return {"up": self.s_up, "compute": self.s_cmp, "blend": self.s_bl}[which]

    def signal(self, which, tables, kind, src, value):
        """Store `value` into counter (kind, src) of every table in `tables` (device pointers, local or peer)."""
        arr = (ctypes.c_void_p * len(tables))(*[t + 4 * flag_index(kind, src) for t in tables])
        with self.torch.cuda.device(self.dev):
            self.N.check(self.lib.innfer_stream_signal(arr, len(tables), int(value) & 0xFFFFFFFF,
                                                       ctypes.c_void_p(self._stream(which).cuda_stream)))

    def wait(self, which, kind, srcs, value):
        """Hold the stream until counters (kind, s) of THIS rank's table are >= value for every s in srcs."""
        if value <= 0 or not srcs:
            return
        arr = (ctypes.c_void_p * len(srcs))(*[self.flags_ptr + 4 * flag_index(kind, s) for s in srcs])
        err = ctypes.c_void_p(self.flags_ptr + 4 * 3 * MAX_RANKS)
        with self.torch.cuda.device(self.dev):
            self.N.check(self.lib.innfer_stream_wait(arr, len(srcs), int(value) & 0xFFFFFFFF, err, self.timeout_ms,
                                                     ctypes.c_void_p(self._stream(which).cuda_stream)))

    def push_frame(self, img, slot, peer_lr_ptrs):
        """Owner: pinned host frame -> own LR slot -> every peer's LR slot (upload stream)."""
        torch = self.torch
        st = ctypes.c_void_p(self.s_up.cuda_stream)
        off = slot * self.lr_bytes
        sl = self._nown & 1
        if isinstance(img, torch.Tensor) and img.is_pinned():
            # a pinned uint8 tensor is copied from where it is (the caller keeps it alive until the frame is done)
            if img.dtype != torch.uint8 or tuple(img.shape) != (self.H, self.W, 3) or not img.is_contiguous():
                raise ValueError("expected a contiguous uint8 [%d,%d,3] frame" % (self.H, self.W))
            src = img.data_ptr()
        else:
            img = np.asarray(img)
            if img.dtype != np.uint8 or img.shape != (self.H, self.W, 3):
                raise ValueError("expected a uint8 [%d,%d,3] frame" % (self.H, self.W))
            # everything here is asynchronous, so the host may run frames ahead of the device: the staging slot is
            # rewritten only after the H2D that last read it has completed
            if self._stage_ev[sl] is not None:
                self._stage_ev[sl].synchronize()
            self.h_in[sl].numpy()[...] = img
            src = self.h_in[sl].data_ptr()
        with torch.cuda.device(self.dev):
            self.N.check(self.lib.innfer_memcpy_async(ctypes.c_void_p(self.lr_ptr + off), ctypes.c_void_p(src),
                                                      self.lr_bytes, st))
            ev = torch.cuda.Event()
            ev.record(self.s_up)
            self._stage_ev[sl] = ev
            for p in peer_lr_ptrs:
                if p != self.lr_ptr:
                    self.N.check(self.lib.innfer_memcpy_async(ctypes.c_void_p(p + off), ctypes.c_void_p(self.lr_ptr + off),
                                                              self.lr_bytes, st))

    def forward_range(self, slot, tiles_ptr, t0, t1):
        """Tiles [t0, t1) of the frame in LR slot `slot` -> tiles_ptr (the owner's buffer, maybe peer memory)."""
        with self.torch.cuda.device(self.dev):
            self.N.check(self.lib.innfer_rrdb_forward_tile_range(
                self.eng._h, ctypes.c_void_p(self.lr_ptr + slot * self.lr_bytes), self.N.INNFER_U8, self.H, self.W,
                self.patch, self.step, t0, t1, ctypes.c_void_p(tiles_ptr), ctypes.c_void_p(self.s_cmp.cuda_stream)))

    def blend(self, frame_index):
        """Owner: gather-blend of the complete tile buffer into the uint8 frame (blend stream)."""
        with self.torch.cuda.device(self.dev):
            self.N.check(self.lib.innfer_rrdb_blend_tiles(self.eng._h, ctypes.c_void_p(self.tiles_ptr), self.H, self.W,
                                                          self.patch, self.step, self.d_out.data_ptr(), self.N.INNFER_U8,
                                                          ctypes.c_void_p(self.s_bl.cuda_stream)))

    def download(self, frame_index):
        """Owner: D2H of the blended frame into a pinned result slot (blend stream); completion event kept."""
        slot = self._nown & 1
        with self.torch.cuda.device(self.dev):
            self.N.check(self.lib.innfer_memcpy_async(ctypes.c_void_p(self.h_out[slot].data_ptr()),
                                                      ctypes.c_void_p(self.d_out.data_ptr()), self.d_out.numel(),
                                                      ctypes.c_void_p(self.s_bl.cuda_stream)))
            ev = self.torch.cuda.Event()
            ev.record(self.s_bl)
        self._done[frame_index] = (ev, slot)
        self._nown += 1

    def fetch(self, frame_index):
        """Owner: block until the frame's D2H has landed; returns the pinned result slot as a numpy view (valid until
        this rank has owned two more frames)."""
        ev, slot = self._done.pop(frame_index)
        ev.synchronize()
        return self.h_out[slot].numpy()

    def errors(self):
        """Number of device-side waits that timed out (synchronises the device)."""
        with self.torch.cuda.device(self.dev):
            self.torch.cuda.synchronize(self.dev)
            host = (ctypes.c_uint32 * 1)()
            self.N.check(self.lib.innfer_memcpy_async(host, ctypes.c_void_p(self.flags_ptr + 4 * 3 * MAX_RANKS), 4, None))
            self.torch.cuda.synchronize(self.dev)
        return int(host[0])

    def synchronize(self):
        self.torch.cuda.synchronize(self.dev)

    def close(self):
        self.synchronize()
        for p in self._opened:
            self.lib.innfer_ipc_close(ctypes.c_void_p(p))
        self._opened = []
        for name in ("lr_ptr", "tiles_ptr", "flags_ptr"):
            p = getattr(self, name, None)
            if p:
                self.lib.innfer_device_free(ctypes.c_void_p(p))
                setattr(self, name, None)


class TileShardedUpscaler:
    """Upscales a stream of frames with all ranks of ``group`` working on every frame (see the module docstring for
    the protocol).  Every rank calls ``submit(f, img)`` for f = 0, 1, 2, ... in order (``img`` is needed on the
    owner only); the owner reads the result with ``result(f)``.  ``upscale`` does both."""

    def __init__(self, backend, dist, group=None):
        self.be, self.dist, self.group = backend, dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > MAX_RANKS:
            raise ValueError("at most %d ranks" % MAX_RANKS)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, backend.export_handles(), group=group)
        # ptrs[r] = [lr_ptr, tiles_ptr, flags_ptr] of rank r as seen from this process
        self.ptrs = [backend.local_ptrs() if r == self.rank else backend.open_handles(gathered[r])
                     for r in range(self.world)]
        self.tables = [p[2] for p in self.ptrs]
        self.ranges = partition(backend.ntiles, self.world)
        self.everyone = list(range(self.world))
        self.next_frame = 0
        dist.barrier(group=group)   # every table is zeroed and mapped before the first signal

    def submit(self, frame_index, img=None):
        if frame_index != self.next_frame:
            raise ValueError("frames must be submitted in order (expected %d)" % self.next_frame)
        self.next_frame += 1
        be, f, G = self.be, frame_index, self.world
        owner, slot, v = frame_owner(f, G), f % 2, f + 1
        if self.rank == owner:
            if img is None:
                raise ValueError("the owner rank must supply the frame")
            be.wait("up", TILES_DONE, self.everyone, f - 1)
            be.push_frame(img, slot, [p[0] for p in self.ptrs])
            be.signal("up", self.tables, LR_READY, owner, v)
        be.wait("compute", LR_READY, [owner], v)
        be.wait("compute", BLEND_DONE, [owner], f - G + 1)
        t0, t1 = self.ranges[self.rank]
        if t1 > t0:
            be.forward_range(slot, self.ptrs[owner][1], t0, t1)
        be.signal("compute", self.tables, TILES_DONE, self.rank, v)
        if self.rank == owner:
            be.wait("blend", TILES_DONE, self.everyone, v)
            be.blend(f)
            be.signal("blend", self.tables, BLEND_DONE, owner, v)
            be.download(f)

    def result(self, frame_index):
        """Owner only: the upscaled uint8 frame (blocks until it has landed on the host)."""
        if frame_owner(frame_index, self.world) != self.rank:
            return None
        return self.be.fetch(frame_index)

    def upscale(self, frame_index, img=None):
        """submit + result: returns the upscaled uint8 image on the owner, None elsewhere."""
        self.submit(frame_index, img)
        return self.result(frame_index)

    def close(self):
        self.be.synchronize()
        self.dist.barrier(group=self.group)   # nobody unmaps while a peer may still store or signal
        nerr = self.be.errors()
        self.be.close()
        if nerr:
            raise RuntimeError("tile-sharded pipeline: %d device-side waits timed out" % nerr)
