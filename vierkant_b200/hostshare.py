"""Host-side plumbing for ONE chain split over several single-GPU worker processes (north_star: "independent
block rows and mip levels shard across the 8 GPUs of one box by simple partitioning, with no NCCL collective, and results
are gathered back over pinned async copies").

The reference hands every 4-block-row batch of a level to a thread pool that writes into one `std::vector<block_t>` per
level (src/texture_block_compression.cpp:107-139).  With one process per GPU the equivalent of that shared vector is a
POSIX shared-memory mapping every worker page-locks (vkt_bcn_cuda_host_register): each GPU's copy engine then writes its
block rows at their final position, and the "gather" is nothing but those copies landing.  This module holds the two
pieces a driver needs around vkt_bcn_cuda_compress_shard_begin / _end: named shared buffers and a flag barrier that moves
no data.  No compute happens here.
"""
from __future__ import annotations

import mmap
import os
import time

import numpy as np


def _candidate_dirs() -> list[str]:
    return [d for d in ("/dev/shm", os.environ.get("TMPDIR", "/tmp"), "/tmp") if os.path.isdir(d) and os.access(d, os.W_OK)]


def _dir_with_room(nbytes: int) -> str:
    """First candidate directory whose file system can hold `nbytes` more (a tmpfs that is too small would only fail later,
    with SIGBUS on the first touch of a page it cannot back)."""
    for d in _candidate_dirs():
        try:
            st = os.statvfs(d)
            if st.f_bavail * st.f_frsize >= nbytes + (64 << 20):
                return d
        except OSError:
            continue
    raise OSError(f"no shared-memory directory with {nbytes >> 20} MB free ({_candidate_dirs()})")


class SharedBuffer:
    """A named byte buffer mapped by every worker of a job.  `create=True` (one worker) makes and sizes it, the others
    attach after the job's start-up barrier.  `array` is a uint8 numpy view; the creator unlinks the name on close()."""

    def __init__(self, name: str, nbytes: int, create: bool):
        self.nbytes = int(nbytes)
        self.owner = create
        if create:
            for d in _candidate_dirs():  # a leftover of a killed run must not be what the other workers find
                try:
                    os.unlink(os.path.join(d, name))
                except OSError:
                    pass
            self.path = os.path.join(_dir_with_room(self.nbytes), name)
        else:  # wherever the creator found room
            found = [os.path.join(d, name) for d in _candidate_dirs() if os.path.exists(os.path.join(d, name))]
            if not found:
                raise FileNotFoundError(f"shared buffer {name} not found in {_candidate_dirs()}")
            self.path = found[0]
        flags = os.O_RDWR | (os.O_CREAT | os.O_TRUNC if create else 0)
        fd = os.open(self.path, flags, 0o600)
        try:
            if create:
                os.ftruncate(fd, max(self.nbytes, 1))
            elif os.fstat(fd).st_size < self.nbytes:
                raise ValueError(f"{self.path}: {os.fstat(fd).st_size} bytes, expected {self.nbytes}")
            self.map = mmap.mmap(fd, max(self.nbytes, 1), mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
        finally:
            os.close(fd)
        self.array = np.frombuffer(self.map, dtype=np.uint8, count=self.nbytes)

    def close(self):
        self.array = None
        try:
            self.map.close()
        except (BufferError, ValueError):
            pass  # a view is still alive somewhere: the mapping goes with the process
        if self.owner:
            try:
                os.unlink(self.path)
            except OSError:
                pass


class FlagBarrier:
    """Barrier over the workers of a job through one cache line per worker in shared memory: worker r publishes the number
    of the phase it has finished, waiters spin until everybody (or one given worker) got there.  What the shard API calls
    "the caller's barrier": it orders the hand-over buffer's writers before worker 0's reads and moves no data."""

    def __init__(self, buf: SharedBuffer, rank: int, world: int):
        assert buf.nbytes >= 64 * world
        self.flags = buf.array[:64 * world].view(np.int64).reshape(world, 8)
        self.rank, self.world = rank, world

    def reset(self):
        self.flags[self.rank, 0] = 0

    def arrive(self, phase: int):
        self.flags[self.rank, 0] = phase

    def wait_all(self, phase: int, timeout_s: float = 180.0):
        t0 = time.monotonic()
        while int(self.flags[:, 0].min()) < phase:
            if time.monotonic() - t0 > timeout_s:
                raise TimeoutError(f"worker {self.rank}: barrier phase {phase} not reached by all workers")

    def wait_for(self, other: int, phase: int, timeout_s: float = 180.0):
        t0 = time.monotonic()
        while int(self.flags[other, 0]) < phase:
            if time.monotonic() - t0 > timeout_s:
                raise TimeoutError(f"worker {self.rank}: worker {other} did not reach phase {phase}")
