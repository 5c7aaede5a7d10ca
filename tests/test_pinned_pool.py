"""CPU tier: ownership semantics of the engine's result pool (``pockit_b200.engine.PinnedPool``).
The reference returns a fresh array per call; the pool must never hand a buffer out
again while any array -- or view of an array -- it returned is still referenced."""
import ctypes as C
import gc

import numpy as np

from pockit_b200.engine import PinnedPool


class FakeLib:
    """pk_alloc_host / pk_free_host stand-ins on the C heap (page-locking needs a CUDA device)."""

    def __init__(self):
        self.libc = C.CDLL(None)
        self.libc.malloc.restype = C.c_void_p
        self.libc.malloc.argtypes = [C.c_size_t]
        self.libc.free.argtypes = [C.c_void_p]
        self.live = set()

    def pk_alloc_host(self, nbytes):
        p = self.libc.malloc(nbytes)
        self.live.add(p)
        return p

    def pk_free_host(self, p):
        self.live.remove(p)
        self.libc.free(p)


def test_buffer_returns_only_when_every_view_is_gone():
    lib = FakeLib()
    pool = PinnedPool(lib)
    a = pool.take(10000)
    addr = a.ctypes.data
    a[:] = 1.0
    view = a[100:200].reshape(10, 10)  # a slice of the result the caller keeps
    del a
    gc.collect()
    b = pool.take(10000)  # the first buffer is still referenced through `view`: a second one is used
    assert b.ctypes.data != addr
    b[:] = 2.0
    assert np.all(view == 1.0)
    del view
    gc.collect()
    c = pool.take(10000)  # now the first buffer is free again
    assert c.ctypes.data == addr and pool.allocations == 2
    del b, c
    gc.collect()
    assert pool.bytes == 2 * 80000
    pool.close()
    assert pool.bytes == 0 and not lib.live


def test_drop_and_ask_again_reuses_one_buffer_and_cap_falls_back_to_pageable():
    lib = FakeLib()
    pool = PinnedPool(lib, max_bytes=3 * 80000, keep_free=1)
    seen = set()
    for _ in range(5):  # a solver that lets go of the previous result before asking for the next
        r = pool.take(10000)
        seen.add(r.ctypes.data)
        del r
        gc.collect()
    assert len(seen) == 1 and pool.allocations == 1
    held = [pool.take(10000) for _ in range(5)]  # a caller that keeps everything: beyond the cap, pageable arrays
    assert pool.allocations == 3 and sum(h.flags.owndata for h in held) == 2
    del held
    gc.collect()
    assert pool.bytes == 80000  # keep_free = 1: the others went back to the allocator
    leased = pool.take(10000)
    pool.close()  # leases outlive the pool; their memory is released when they die
    assert leased.ctypes.data in lib.live
    leased[:] = 3.0
    del leased
    gc.collect()
    assert not lib.live
