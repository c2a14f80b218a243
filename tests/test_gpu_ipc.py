"""rr_ipc_*: a device buffer of one process as the output buffer of a chain in another (the gather of a channelizer's
outputs without a collective kernel).  Two processes on the same GPU here; across GPUs the same calls enable peer access."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _child(handle_bytes, cap, n_streams, q):
    try:
        sys.path.insert(0, ROOT)
        import ctypes as C

        import torch

        import radiorust_b200 as rr
        from oracle import radiorust_oracle as orc
        from radiorust_b200 import _ffi

        lib = _ffi.load()
        ctx = rr.Context(0)
        ptr = C.c_void_p()
        hb = (C.c_ubyte * 64).from_buffer_copy(handle_bytes)
        _ffi.check(lib.rr_ipc_open(ctx._h, hb, C.byref(ptr)))
        sr, n, k = 48000.0, 512, 4
        x = np.stack([orc.synth_noise(50 + s, k * n, "f32") for s in range(n_streams)])
        xd = torch.from_numpy(x.view(np.float32).reshape(n_streams, k * n, 2)).cuda()
        ch = rr.Chain(ctx, [rr.GainControl(0.5)], "f32", n_streams=n_streams)
        cnt, _ = ch.push_device(sr, n, k, xd.data_ptr(), k * n, ptr.value, cap, cap)
        ch.sync()
        ch.close()
        _ffi.check(lib.rr_ipc_close(ctx._h, ptr))
        ctx.close()
        q.put(("ok", int(cnt)))
    except Exception as e:  # pragma: no cover
        q.put(("error", repr(e)))


def test_chain_output_lands_in_another_process_buffer():
    import ctypes as C

    import radiorust_b200 as rr
    from oracle import radiorust_oracle as orc
    from radiorust_b200 import _ffi

    lib = _ffi.load()
    ctx = rr.Context(0)
    S, cap = 3, 2048
    ptr = C.c_void_p()
    _ffi.check(lib.rr_device_alloc(ctx._h, S * cap * 8, C.byref(ptr)))
    zeros = np.zeros((S, cap), dtype=np.complex64)
    _ffi.check(lib.rr_memcpy_h2d(ctx._h, ptr, zeros.ctypes.data_as(C.c_void_p), zeros.nbytes))
    handle = (C.c_ubyte * 64)()
    _ffi.check(lib.rr_ipc_export(ctx._h, ptr, handle))
    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    p = mpctx.Process(target=_child, args=(bytes(handle), cap, S, q))
    p.start()
    status, val = q.get(timeout=300)
    p.join(timeout=60)
    assert status == "ok", val
    assert val == 2048
    got = np.empty((S, cap), dtype=np.complex64)
    _ffi.check(lib.rr_memcpy_d2h(ctx._h, got.ctypes.data_as(C.c_void_p), ptr, got.nbytes))
    for s in range(S):
        want = orc.synth_noise(50 + s, 4 * 512, "f32") * np.float32(0.5)
        assert np.array_equal(got[s, :2048], want)
    lib.rr_device_free(ctx._h, ptr)
    ctx.close()
