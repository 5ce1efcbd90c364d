"""Multi-GPU sharding of one frame: row-cyclic partition + peer-mapped output image.

New work (the reference is single-device).  Pixels are independent given the replicated read-only
textures (~43 MB per GPU) and 328 bytes of uniforms, so the frame shards with NO data-path collective:
rank r marches row blocks b with b % world == r (blocks of `row_block` rows; contiguous bands would be
badly unbalanced because rows below the horizon are free) and its kernel stores finished
pixels straight into rank 0's image, mapped through CUDA IPC over NVLink.  torch.distributed is only the
plumbing that carries the 64-byte handle and the barriers.
"""
import numpy as np


def owned_rows(H, rank, world, row_block, snake=False):
    """Rows marched by `rank` -- the same enumeration as the kernel (csrc/common.h: owned_block) and mm_dispatch.
    snake (MM_ROWS_SNAKE): odd rounds of the cyclic assignment run in reverse rank order, so no rank is systematically
    nearer the (expensive) horizon inside every round."""
    rows = []
    nblocks = (H + row_block - 1) // row_block
    k = 0
    while True:
        b = k * world + ((world - 1 - rank) if (snake and k % 2) else rank)
        if b >= nblocks:
            break
        rows.extend(y for y in range(b * row_block, min(H, (b + 1) * row_block)))
        k += 1
    return np.asarray(rows, dtype=np.int64)


def partition_is_exact_cover(H, world, row_block, snake=False):
    seen = np.zeros(H, np.int32)
    for r in range(world):
        seen[owned_rows(H, r, world, row_block, snake)] += 1
    return bool((seen == 1).all())


def exchange_handle(handle, rank, world, dist=None, src=0):
    """Broadcast rank `src`'s 64-byte IPC handle to every rank (works on gloo and nccl process groups)."""
    if world == 1:
        return handle
    box = [handle if rank == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


class SharedFrame:
    """The output image of a sharded frame: allocated on rank 0, mapped on every other rank."""

    def __init__(self, cs, rank, world, dist=None):
        self.cs, self.rank, self.world = cs, rank, world
        self.remote_ptr = None
        handle = None
        if rank == 0:
            self.ptr, self.pitch = cs.allocOutput()
            if world > 1:
                handle = cs.ipcGetHandle(self.ptr)
        handle = exchange_handle(handle, rank, world, dist)
        if rank != 0:
            self.remote_ptr = cs.ipcOpenHandle(handle)
            self.ptr, self.pitch = self.remote_ptr, cs.width * 16
            cs.bindOutput(self.ptr, self.pitch)

    def close(self):
        if self.remote_ptr is not None:
            self.cs.ipcCloseHandle(self.remote_ptr)
            self.remote_ptr = None


class HostBarrier:
    """Barrier between the processes of one node through shared memory: one int64 epoch per rank, a cache line apart, in a buffer
    every process has mapped (at least world * 64 bytes, zero-initialised).  wait() publishes this rank's next epoch and spins until
    every rank has published it.  x86 stores are ordered, and the ranks only ever increase their own slot."""

    def __init__(self, buffer, rank, world, offset=0, timeout_s=120.0):
        self.rank = rank
        self._epochs = np.frombuffer(buffer, dtype=np.int64, count=world * 8, offset=offset)[::8]
        self._epoch = 0
        self._timeout_s = timeout_s

    def wait(self):
        """Raises RuntimeError when a rank has not arrived within `timeout_s` (a peer process died): a sharded frame must fail, not hang."""
        import time
        self._epoch += 1
        self._epochs[self.rank] = self._epoch
        e = self._epochs
        spins, deadline = 0, None
        while int(e.min()) < self._epoch:
            spins += 1
            if (spins & 0xfff) == 0:                  # look at the clock every 4096 polls only: the common wait is microseconds
                now = time.monotonic()
                if deadline is None:
                    deadline = now + self._timeout_s
                elif now > deadline:
                    late = [r for r, v in enumerate(e.tolist()) if v < self._epoch]
                    raise RuntimeError(f"HostBarrier: ranks {late} did not reach epoch {self._epoch} within {self._timeout_s:.0f} s")

    def release(self):
        self._epochs = None


class SharedHostFrame:
    """The HOST copy of a sharded frame: one POSIX shared-memory mapping opened by every rank's process, page-locked and
    mapped into each rank's GPU.  Every rank's march kernel stores its pixels there as it finishes them (next to the store
    into rank 0's device image), so the device->host transfer of the frame runs over N PCIe links in parallel and overlaps
    the march; after a barrier the frame is complete in `self.array` on every rank -- no gather, no copy.
    The same mapping carries one epoch counter per rank behind the frame: `barrier()` is a host-side barrier between the ranks'
    processes through that shared memory (a store and a spin on 8 cache lines), for the per-frame "every rank's stream has drained"
    rendezvous -- no collective kernel, no device-wide synchronize."""

    def __init__(self, cs, rank, world, dist=None):
        import mmap
        import os
        self.cs, self.rank, self.world = cs, rank, world
        frame_bytes = cs.width * cs.height * 16
        nbytes = frame_bytes + 4096
        name = [f"/dev/shm/marshmallow_frame_{os.getpid()}" if rank == 0 else None]
        if rank == 0:
            with open(name[0], "wb") as f:
                f.truncate(nbytes)
        if world > 1:
            dist.broadcast_object_list(name, src=0)
        self.path = name[0]
        self._f = open(self.path, "r+b")
        self._map = mmap.mmap(self._f.fileno(), nbytes)
        self.array = np.frombuffer(self._map, dtype=np.float32, count=cs.width * cs.height * 4).reshape(cs.height, cs.width, 4)
        self._barrier = HostBarrier(self._map, rank, world, offset=frame_bytes)
        cs.hostRegister(self.array)
        cs.bindHostMirror(self.array)
        if world > 1:
            dist.barrier()                    # every rank has the file open: rank 0 may unlink the name now
        if rank == 0:
            os.unlink(self.path)

    def barrier(self):
        """host-side barrier of the ranks' processes (each must have drained its own stream first)"""
        self._barrier.wait()

    def close(self):
        if self.array is not None:
            self.cs.bindHostMirror(None)
            self.cs.hostUnregister(self.array)
            self.array = None
            self._barrier.release()
            try:
                self._map.close()
            except BufferError:           # a caller still holds a view of the frame; the mapping goes with its last reference
                pass
            self._f.close()


class FrameRing:
    """Frame-parallel animation (BASELINE config 5): frame k is rendered whole by rank k % world.  Rank 0 owns one
    image slot per rank; every rank binds ITS slot (mapped through CUDA IPC on ranks != 0) and stores finished frames
    there over NVLink, so rank 0 always holds the latest frame of every rank without a gather step."""

    def __init__(self, cs, rank, world, dist=None):
        self.cs, self.rank, self.world = cs, rank, world
        self.frame_bytes = cs.width * cs.height * 16
        self.base = self.remote = None
        handle = None
        if rank == 0:
            self.base = cs.allocDevice(self.frame_bytes * world)
            if world > 1:
                handle = cs.ipcGetHandle(self.base)
        handle = exchange_handle(handle, rank, world, dist)
        if rank != 0:
            self.remote = cs.ipcOpenHandle(handle)
        self.slot = (self.base if rank == 0 else self.remote) + rank * self.frame_bytes
        cs.bindOutput(self.slot, cs.width * 16)

    def frames_of(self, n_frames):
        return list(range(self.rank, n_frames, self.world))

    def close(self):
        if self.remote is not None:
            self.cs.ipcCloseHandle(self.remote)
            self.remote = None
        if self.base is not None:
            self.cs.freeDevice(self.base)
            self.base = None
