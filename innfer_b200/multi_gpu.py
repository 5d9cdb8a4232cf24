"""Multi-GPU execution of the RRDB path: one process per GPU, no collective on the data path.

The path shards trivially: tiles of one frame, and frames, are independent.
Two modes:

* image sharding (``frames_for_rank``): rank r upscales frames r, r+G, ... on its own GPU --
  throughput mode, used by bench.py (weak scaling);
* tile sharding (``TileShardedUpscaler``): the ranks split the row-major tile list of ONE frame
  into contiguous ranges; every rank's last conv stores its finished tiles straight into the frame
  owner's tile buffer through a CUDA-IPC peer mapping (NVLink/NVSwitch P2P stores), then the owner
  runs the single gather-blend kernel in fixed tile order, so the result is bit-identical for any
  number of ranks.  torch.distributed is used for the control plane only (handle exchange and two
  barriers per frame).
"""
import ctypes

import numpy as np


def partition(n_items, world):
    """Contiguous ranges of ceil(n/world) items: [(begin, end)] * world."""
    per = -(-n_items // world)
    return [(min(r * per, n_items), min((r + 1) * per, n_items)) for r in range(world)]


def frame_owner(frame_index, world):
    return frame_index % world


def frames_for_rank(n_frames, rank, world):
    return list(range(rank, n_frames, world))


class NativeTileBackend:
    """Device-side operations of the tile-sharded mode on top of the C-ABI."""

    def __init__(self, engine, H, W, patch=200, step=0.5):
        from . import _native as N
        self.N, self.lib, self.eng = N, N.load(), engine
        self.H, self.W, self.patch, self.step = H, W, patch, step
        self.scale = engine.cfg["scale"]
        self._opened = []
        ptr, nbytes, tb = ctypes.c_void_p(), ctypes.c_uint64(), ctypes.c_uint64()
        N.check(self.lib.innfer_rrdb_tile_buffer(engine._h, H, W, patch, step, ctypes.byref(ptr), ctypes.byref(nbytes),
                                                 ctypes.byref(tb)))
        self.tiles_ptr = ptr.value
        lr = ctypes.c_void_p()
        N.check(self.lib.innfer_device_alloc(engine.index, H * W * 3, ctypes.byref(lr)))
        self.lr_ptr = lr.value
        n, ts = ctypes.c_int(), ctypes.c_int()
        N.check(self.lib.innfer_tiles_plan(H, W, patch, step, None, 0, ctypes.byref(n), ctypes.byref(ts)))
        self.ntiles = n.value

    def export_handles(self):
        out = []
        for p in (self.lr_ptr, self.tiles_ptr):
            buf = (ctypes.c_uint8 * 64)()
            self.N.check(self.lib.innfer_ipc_export(ctypes.c_void_p(p), buf))
            out.append(bytes(buf))
        return out

    def open_handles(self, handles):
        ptrs = []
        for hd in handles:
            buf = (ctypes.c_uint8 * 64).from_buffer_copy(hd)
            p = ctypes.c_void_p()
            self.N.check(self.lib.innfer_ipc_open(buf, ctypes.byref(p)))
            self._opened.append(p.value)
            ptrs.append(p.value)
        return ptrs

    def local_ptrs(self):
        return [self.lr_ptr, self.tiles_ptr]

    def upload(self, img):
        import torch
        img = np.ascontiguousarray(img)
        if img.dtype != np.uint8 or img.shape != (self.H, self.W, 3):
            raise ValueError("expected a uint8 [%d,%d,3] frame" % (self.H, self.W))
        with torch.cuda.device(self.eng.index):
            self.N.check(self.lib.innfer_device_upload(ctypes.c_void_p(self.lr_ptr), img.ctypes.data, img.nbytes, None))

    def forward_range(self, lr_ptr, tiles_ptr, t0, t1):
        import torch
        with torch.cuda.device(self.eng.index):
            stream = torch.cuda.current_stream().cuda_stream
            self.N.check(self.lib.innfer_rrdb_forward_tile_range(self.eng._h, ctypes.c_void_p(lr_ptr), self.N.INNFER_U8,
                                                                 self.H, self.W, self.patch, self.step, t0, t1,
                                                                 ctypes.c_void_p(tiles_ptr), ctypes.c_void_p(stream)))
            torch.cuda.current_stream().synchronize()

    def blend(self, tiles_ptr):
        import torch
        with torch.cuda.device(self.eng.index):
            out = torch.empty((self.scale * self.H, self.scale * self.W, 3), dtype=torch.uint8,
                              device=torch.device("cuda", self.eng.index))
            stream = torch.cuda.current_stream().cuda_stream
            self.N.check(self.lib.innfer_rrdb_blend_tiles(self.eng._h, ctypes.c_void_p(tiles_ptr), self.H, self.W, self.patch,
                                                          self.step, out.data_ptr(), self.N.INNFER_U8,
                                                          ctypes.c_void_p(stream)))
            return out.cpu().numpy()

    def close(self):
        for p in self._opened:
            self.lib.innfer_ipc_close(ctypes.c_void_p(p))
        self._opened = []
        if self.lr_ptr:
            self.lib.innfer_device_free(ctypes.c_void_p(self.lr_ptr))
            self.lr_ptr = None


class TileShardedUpscaler:
    """Upscales frames one at a time with all ranks of ``group`` working on each frame."""

    def __init__(self, backend, dist, group=None):
        self.be, self.dist, self.group = backend, dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        gathered = [None] * self.world
        dist.all_gather_object(gathered, backend.export_handles(), group=group)
        # ptrs[r] = [lr_ptr, tiles_ptr] of rank r as seen from this process
        self.ptrs = [backend.local_ptrs() if r == self.rank else backend.open_handles(gathered[r])
                     for r in range(self.world)]
        self.ranges = partition(backend.ntiles, self.world)

    def upscale(self, frame_index, img=None):
        """All ranks call this for every frame; ``img`` (uint8 HWC BGR) is needed on the owner only.
        Returns the upscaled uint8 image on the owner, None elsewhere."""
        owner = frame_owner(frame_index, self.world)
        if self.rank == owner:
            if img is None:
                raise ValueError("the owner rank must supply the frame")
            self.be.upload(img)
        self.dist.barrier(group=self.group)          # frame is resident on the owner
        lr_ptr, tiles_ptr = self.ptrs[owner]
        t0, t1 = self.ranges[self.rank]
        if t1 > t0:
            self.be.forward_range(lr_ptr, tiles_ptr, t0, t1)
        self.dist.barrier(group=self.group)          # every tile landed in the owner's buffer
        return self.be.blend(tiles_ptr) if self.rank == owner else None

    def close(self):
        self.dist.barrier(group=self.group)
        self.be.close()
