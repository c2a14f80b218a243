"""The pinned chunk pool at the chain edges (rr_pool_*; bufferpool.rs:187-222) and the double-buffered host push."""
import numpy as np
import pytest

from oracle import radiorust_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import radiorust_b200 as rr

    c = rr.Context(0)
    yield c
    c.close()


def test_pool_recycles_oldest_first(ctx):
    import radiorust_b200 as rr

    pool = rr.PinnedChunkBufPool(ctx)
    a = pool.get((4096,), np.complex64)
    b = pool.get((4096,), np.complex64)
    pa, pb = a.ctypes.data, b.ctypes.data
    assert pa != pb
    a[:] = 1.0
    pool.put(a)
    pool.put(b)
    assert pool.stats() == (2, 0, 2, 0)
    c = pool.get((1000,), np.complex64)  # the oldest recycled buffer comes back first (FIFO recycler channel)
    assert c.ctypes.data == pa
    d = pool.get((4096,), np.complex64)
    assert d.ctypes.data == pb
    assert pool.stats() == (2, 2, 0, 2)
    pool.put(c)
    e = pool.get((1 << 16,), np.complex64)  # too small to reuse: replaced by a new allocation
    assert e.shape == (1 << 16,)
    assert pool.stats()[0] == 3
    pool.put(d)
    pool.put(e)
    pool.trim()
    assert pool.stats()[2:] == (0, 0)
    pool.close()


def test_back_to_back_host_pushes_from_pooled_chunks(ctx):
    """Pushes are enqueued without a sync in between (two staging slots, separate copy streams): the outputs must
    be those of the same pushes done one by one."""
    import radiorust_b200 as rr

    sr, n, S, k, pushes = 2_400_000.0, 4096, 3, 4, 6
    stages = [rr.FreqShifter(0.0), rr.Filter.new(orc.lowpass(3000.0)), rr.Downsampler(64, 48000.0, 6000.0)]
    shifts = [-577000.0, 123457.0, 0.0]
    x = np.stack([orc.synth_noise(4100 + s, pushes * k * n, "f32") for s in range(S)])
    ref = rr.Chain(ctx, stages, "f32", n_streams=S)
    ref.set_shifts(0, shifts)
    want = []
    for p in range(pushes):
        y, _ = ref.push(sr, np.ascontiguousarray(x[:, p * k * n : (p + 1) * k * n]), n)
        want.append(y.copy())
    ref.close()

    pool = rr.PinnedChunkBufPool(ctx)
    ch = rr.Chain(ctx, stages, "f32", n_streams=S)
    ch.set_shifts(0, shifts)
    ins, outs, counts = [], [], []
    for p in range(pushes):
        xi = pool.get((S, k * n), np.complex64)
        xi[:] = x[:, p * k * n : (p + 1) * k * n]
        cap = max(ch.max_output(sr, n, k), 1)
        yo = pool.get((S, cap), np.complex64)
        cnt, rate = ch.push_host_async(sr, n, k, xi.ctypes.data, k * n, yo.ctypes.data, cap, cap)
        ins.append(xi)
        outs.append(yo)
        counts.append(cnt)
    ch.sync()
    for p in range(pushes):
        assert counts[p] == want[p].shape[1]
        assert np.array_equal(outs[p][:, : counts[p]], want[p])
        pool.put(ins[p])
        pool.put(outs[p])
    assert pool.stats()[3] == 0
    ch.close()
    pool.close()


def test_refused_push_leaves_the_chain_untouched(ctx):
    """A Filter chunk length the device path cannot take is refused before any stage advances (the stages in front
    of the Filter included)."""
    import radiorust_b200 as rr

    sr = 48000.0
    stages = [rr.Rechunker(100), rr.FreqShifter(1000.0), rr.Filter.new(orc.lowpass(3000.0))]
    ch = rr.Chain(ctx, stages, "f32")
    x = orc.synth_noise(5, 4096, "f32")
    ch2 = rr.Chain(ctx, [rr.FreqShifter(1000.0), rr.Filter.new(orc.lowpass(3000.0))], "f32")
    y_ok, _ = ch2.push(sr, x, 1024)
    big = (1 << 25) + 2  # beyond every plan
    with pytest.raises(rr.RadiorustError) as ei:
        ch2.push_device(sr, big, 2, 16, big * 2, 16, 0, 1)
    assert ei.value.code == -3
    y2, _ = ch2.push(sr, x, 1024)  # same shape as before the refusal: no redesign, history intact
    oc = orc.Chain([orc.FreqShifter("f32", 1.0, 1000.0), orc.Filter.new("f32", orc.lowpass(3000.0))])
    w1 = oc.run(sr, x, 1024)
    w2 = oc.run(sr, x, 1024)
    assert y_ok.shape[1] == len(w1) and y2.shape[1] == len(w2) == 4096
    assert orc.rel_l2(y2[0], w2) <= 1e-5
    ch.close()
    ch2.close()
