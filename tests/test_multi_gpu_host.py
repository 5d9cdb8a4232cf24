"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: tile partition, frame ownership and
the TileShardedUpscaler control flow, with a shared-memory stand-in for the CUDA-IPC buffers."""
import multiprocessing as mp
import os
import time
from multiprocessing import shared_memory

import numpy as np
import pytest
import torch

from innfer_b200 import multi_gpu as MG
from oracle import rrdb_oracle as O


def test_partition_and_ownership():
    assert MG.partition(190, 8) == [(0, 24), (24, 48), (48, 72), (72, 96), (96, 120), (120, 144), (144, 168), (168, 190)]
    assert MG.partition(190, 1) == [(0, 190)]
    assert MG.partition(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]
    for n, w in ((190, 2), (84, 4), (25, 8), (1, 8)):
        parts = MG.partition(n, w)
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    assert [MG.frame_owner(f, 4) for f in range(6)] == [0, 1, 2, 3, 0, 1]
    assert MG.frames_for_rank(10, 1, 4) == [1, 5, 9]


H, W, P, S = 40, 56, 32, 2


def _fake_forward(tile):
    """Deterministic stand-in for the network on one [3,p,p] tile -> [3,S*p,S*p]."""
    up = torch.nn.functional.interpolate(tile[None], scale_factor=float(S), mode="nearest")[0]
    return up * 0.5 + 0.25 * up.flip(0)


class ShmBackend:
    """Same interface as NativeTileBackend with multiprocessing shared memory instead of CUDA IPC and synchronous
    "streams": a signal is a store into the peers' tables, a wait polls this rank's own table.  Executing every
    enqueue immediately is one valid serialisation of the stream-ordered protocol, so a deadlock or a missing
    dependency in TileShardedUpscaler shows up here as a timeout or a wrong image."""

    def __init__(self, rank):
        self.ys, self.xs = O.tile_origins(H, P), O.tile_origins(W, P)
        self.ntiles = len(self.ys) * len(self.xs)
        self.tile_elems = 3 * (S * P) * (S * P)
        self.lr_bytes = H * W * 3
        self.lr = shared_memory.SharedMemory(create=True, size=2 * self.lr_bytes)
        self.tiles = shared_memory.SharedMemory(create=True, size=self.ntiles * self.tile_elems * 4)
        self.flags = shared_memory.SharedMemory(create=True, size=(3 * MG.MAX_RANKS + 1) * 4)
        np.ndarray((3 * MG.MAX_RANKS + 1,), np.uint32, self.flags.buf)[:] = 0
        self.opened = []
        self.results = {}
        self.log = []

    def export_handles(self):
        return [self.lr.name, self.tiles.name, self.flags.name]

    def open_handles(self, names):
        shms = [shared_memory.SharedMemory(name=n) for n in names]
        self.opened.extend(shms)
        return shms

    def local_ptrs(self):
        return [self.lr, self.tiles, self.flags]

    def signal(self, which, tables, kind, src, value):
        for t in tables:
            np.ndarray((3 * MG.MAX_RANKS + 1,), np.uint32, t.buf)[MG.flag_index(kind, src)] = value
        self.log.append(("signal", which, kind, src, value))

    def wait(self, which, kind, srcs, value):
        if value <= 0 or not srcs:
            return
        tab = np.ndarray((3 * MG.MAX_RANKS + 1,), np.uint32, self.flags.buf)
        t0 = time.time()
        while any(int(tab[MG.flag_index(kind, s)]) < value for s in srcs):
            if time.time() - t0 > 60:
                raise TimeoutError("wait(%s, kind %d, %s, %d) timed out" % (which, kind, srcs, value))
            time.sleep(0.0005)
        self.log.append(("wait", which, kind, tuple(srcs), value))

    def push_frame(self, img, slot, peer_lrs):
        for lr in peer_lrs:
            np.ndarray((2, H, W, 3), np.uint8, lr.buf)[slot] = img

    def forward_range(self, slot, tiles, t0, t1):
        img = np.ndarray((2, H, W, 3), np.uint8, self.lr.buf)[slot]
        x = O.np2tensor(img)
        out = np.ndarray((self.ntiles, 3, S * P, S * P), np.float32, tiles.buf)
        for t in range(t0, t1):
            y0, x0 = self.ys[t // len(self.xs)], self.xs[t % len(self.xs)]
            out[t] = _fake_forward(x[0, :, y0:y0 + P, x0:x0 + P]).numpy()

    def blend(self, f):
        t = torch.from_numpy(np.ndarray((self.ntiles, 3, S * P, S * P), np.float32, self.tiles.buf).copy())
        self._blended = O.tensor2np(O.recompose(t, H, W, 0.5, S))

    def download(self, f):
        self.results[f] = self._blended

    def fetch(self, f):
        return self.results.pop(f)

    def errors(self):
        return 0

    def synchronize(self):
        pass

    def close(self):
        for s in self.opened:
            s.close()
        for s in (self.lr, self.tiles, self.flags):
            s.close()
            s.unlink()


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    be = ShmBackend(rank)
    up = MG.TileShardedUpscaler(be, dist)
    outs = {}
    nframes = 5
    # frames are submitted ahead of reading the results, as bench.py does (results lag one owned frame behind)
    for f in range(nframes):
        img = np.random.default_rng(100 + f).integers(0, 256, (H, W, 3), dtype=np.uint8)
        up.submit(f, img if MG.frame_owner(f, world) == rank else None)
        g = f - world
        if g >= 0 and MG.frame_owner(g, world) == rank:
            outs[g] = up.result(g).copy()
    for g in range(max(0, nframes - world), nframes):
        if MG.frame_owner(g, world) == rank:
            outs[g] = up.result(g).copy()
    # protocol facts visible in this rank's own log: the tile-buffer back-pressure wait appears from frame `world` on,
    # and every frame this rank owns signalled LR_READY before and BLEND_DONE after its TILES_DONE wait
    waits = [e for e in be.log if e[0] == "wait"]
    assert any(e[2] == MG.BLEND_DONE for e in waits) == (nframes > world)
    for f in range(nframes):
        if MG.frame_owner(f, world) == rank:
            i_lr = be.log.index(("signal", "up", MG.LR_READY, rank, f + 1))
            i_td = be.log.index(("wait", "blend", MG.TILES_DONE, tuple(range(world)), f + 1))
            i_bd = be.log.index(("signal", "blend", MG.BLEND_DONE, rank, f + 1))
            assert i_lr < i_td < i_bd
    up.close()
    q.put((rank, outs))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_tile_sharded_control_flow_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in procs:
        rank, outs = q.get(timeout=120)
        for f, img in outs.items():
            assert MG.frame_owner(f, world) == rank
            got[f] = img
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(got) == [0, 1, 2, 3, 4]
    # single-process reference with the same fake network: must be bit-identical
    for f in range(5):
        img = np.random.default_rng(100 + f).integers(0, 256, (H, W, 3), dtype=np.uint8)
        x = O.np2tensor(img)
        patches, _, _ = O.extract_patches(x, P, 0.5)
        tiles = torch.stack([_fake_forward(patches[i]) for i in range(patches.shape[0])], 0)
        want = O.tensor2np(O.recompose(tiles, H, W, 0.5, S))
        assert np.array_equal(got[f], want)
