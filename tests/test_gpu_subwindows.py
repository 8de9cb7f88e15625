"""Sub-windows: a handle created with sub_windows = K cuts a long decode into K time shards on ITS OWN device (internal
handles, enqueued back to back, carries stitched in C -- ookd_multi.cpp) so that the tail of sub-window j overlaps the
screening of sub-window j+1.  Everything a caller sees must be what the undivided decode gives, i.e. what the reference's
loop (src/ookiedokie.c:238-290) prints: messages, transitions, exit state, also across chained windows and resolve."""
import numpy as np
import pytest

from oracle import oracle as O
from ookiedokie_b200 import binding as B
from ookiedokie_b200 import shard as S
import ookd_testutil as util

pytestmark = pytest.mark.gpu

CASES = [
    ("p3l-nexa2012", "fs32_fs4", 8192, 0.03, 2),
    ("p3l-nexa2012", "fs32_fs4", 8192, 0.03, 4),
    ("unknown-remote1", "fs128_fs16_dec4", 1001, 0.03, 3),
    ("p3l-nexa2012", "fs64_fs8", 4096, 0.01, 5),
]


def _setup(devname, filt, sigma, n_msgs=9, glitch=True):
    dev = O.load_device(devname)
    fields = util.nexa_fields if "nexa" in devname else util.remote_fields
    glitches = ((9000, 100),) if glitch and filt != "fs64_fs8" else ()
    iq, msgs, _ = util.capture(dev, n_msgs, sigma=sigma, amplitude=0.8, phase=0.4, seed=21, fields=fields, glitches=glitches)
    stages = O.load_filter(filt)
    return dev, iq, stages, util.sm_spec(dev, stages)


@pytest.mark.parametrize("devname,filt,spb,sigma,k", CASES)
def test_subwindow_decode_matches_oracle(devname, filt, spb, sigma, k):
    dev, iq, stages, sm = _setup(devname, filt, sigma)
    ref = O.rx(iq, stages, dev, samples_per_buffer=spb)
    assert len(ref["msgs"]) >= 3
    plain = B.Gpu(filter_stages=stages, sm=sm, samples_per_buffer=spb, sm_chunk_buffers=5)
    sub = B.Gpu(filter_stages=stages, sm=sm, samples_per_buffer=spb, sm_chunk_buffers=5, sub_windows=k)
    want = plain.decode(iq)
    assert want["msgs"] == ref["msgs"]
    for rep in range(2):                              # (the second decode replays the internal handles' graphs)
        got = sub.decode(iq)
        assert got["msgs"] == ref["msgs"], (k, rep)
        assert got["n_out"] == ref["n_out"] and got["n_buffers"] == want["n_buffers"] and got["n_in"] == want["n_in"]
        fb, edges = sub.edges()
        assert fb == ref["first_bit"] and np.array_equal(edges, ref["edges"])
        assert got["n_edges"] >= len(ref["edges"])     # (shards count the transitions of their warm-up history too)
        with pytest.raises(B.OokdError):               # the decisions are spread over the internal handles
            sub.bits()
    # device-resident input: one pointer, cut inside the library
    import torch
    t = torch.from_numpy(np.ascontiguousarray(iq).reshape(-1).copy()).cuda()
    got_d = sub.decode((t.data_ptr(), len(iq)))
    assert got_d["msgs"] == ref["msgs"]
    assert np.array_equal(sub.edges()[1], ref["edges"])
    plain.close()
    sub.close()


def test_short_decode_is_not_cut():
    """A capture shorter than K x the history a shard reads goes through the handle itself: every accessor works."""
    dev, iq, stages, sm = _setup("p3l-nexa2012", "fs32_fs4", 0.0, n_msgs=1, glitch=False)
    ref = O.rx(iq, stages, dev, samples_per_buffer=8192, want_bits=True)
    sub = B.Gpu(filter_stages=stages, sm=sm, samples_per_buffer=8192, sub_windows=4)      # default chunks: 64 buffers of history
    assert sub.halo > len(iq) // 4
    got = sub.decode(iq)
    assert got["msgs"] == ref["msgs"]
    assert np.array_equal(sub.bits(), ref["bits"])
    assert np.array_equal(sub.edges()[1], ref["edges"])
    sub.close()


@pytest.mark.parametrize("devname,filt,spb,sigma,k", CASES[:3])
def test_chained_windows_and_resolve(devname, filt, spb, sigma, k):
    """Two consecutive windows through a sub-window handle: (a) the second entered with the first one's exit, (b) the
    second begun BEFORE the first has ended (entry from warm-up history) and corrected with resolve if need be, (c) entered
    from a wrong state and resolved."""
    dev, iq, stages, sm = _setup(devname, filt, sigma, n_msgs=14)
    ref = O.rx(iq, stages, dev, samples_per_buffer=spb)
    a = B.Gpu(filter_stages=stages, sm=sm, samples_per_buffer=spb, sm_chunk_buffers=5, sub_windows=k)
    b = B.Gpu(filter_stages=stages, sm=sm, samples_per_buffer=spb, sm_chunk_buffers=5, sub_windows=k)
    halo = a.halo
    align = int(np.lcm(spb, O.filter_total_decimation(stages)))
    n = len(iq)
    cut = max(align, (n // 2) // align * align)
    assert cut > halo
    dec = a.total_decimation
    # (a)
    r1, c1 = a.decode_shard(iq[:cut], 0, cut, False)
    e1 = a.edges()[1]
    r2, c2 = a.decode_shard(iq[cut - halo:], cut, n - cut, True, c1)
    e2 = a.edges()[1]
    assert r1["msgs"] + r2["msgs"] == ref["msgs"]
    assert np.array_equal(np.concatenate([e1, e2]), ref["edges"])
    # (b)
    a.decode_begin(iq[:cut], 0, cut, False)
    b.decode_begin(iq[cut - halo:], cut, n - cut, True)
    q1, d1 = a.decode_end()
    q2, d2 = b.decode_end()
    assert d1 == c1 and q1["msgs"] == r1["msgs"]
    if q2["entry_used"] != d1:
        q2, d2 = b.resolve(d1)
    assert d2 == c2 and q2["msgs"] == r2["msgs"]
    # (c)
    w, _ = b.decode_shard(iq[cut - halo:], cut, n - cut, True, S.INITIAL_CARRY)
    w2, d3 = b.resolve(c1)
    assert d3 == c2 and w2["msgs"] == r2["msgs"]
    assert np.array_equal(b.edges()[1], e2)
    a.close()
    b.close()


def test_subwindows_at_scale_equal_the_undivided_decode():
    """2^27 samples of the benchmark recipe: K = 4 against the plain handle (which test_gpu_scale_parity pins to the
    oracle at this size): same messages, transitions and exit state."""
    import torch
    import bench
    from ookiedokie_b200 import host as H
    n = 1 << 27
    fir = H.Fir(bench.FILTER_NAME)
    dev = H.Device(bench.DEVICE_NAME, bench.FS // fir.total_decimation)
    tog, _ = bench.build_toggles(dev, n)
    i_on, q_on = bench.on_level()
    d = torch.empty(n * 2, dtype=torch.int16, device="cuda")
    B.synth(n, tog, i_on, q_on, bench.noise_scale(), bench.SEED, device_id=0, device_ptr=d.data_ptr(), noise_terms=bench.NOISE_TERMS)
    torch.cuda.synchronize()
    kw = dict(filter_stages=fir.stages, sm=dev.sm_spec(), threshold=bench.THR, samples_per_buffer=bench.SPB)
    plain, sub = B.Gpu(**kw), B.Gpu(sub_windows=4, **kw)
    want, wexit = plain.decode_shard((d.data_ptr(), n), 0, n, True)
    wedges = plain.edges()[1]
    for rep in range(2):
        got, gexit = sub.decode_shard((d.data_ptr(), n), 0, n, True)
        assert gexit == wexit
        assert got["msgs_raw"].tobytes() == want["msgs_raw"].tobytes() and len(want["msgs_raw"]) > 200
        assert np.array_equal(sub.edges()[1], wedges)
    plain.close()
    sub.close()


@pytest.mark.parametrize("chunk_buffers", [8, 16])
def test_silence_and_bursts_longer_than_a_chunk_resolve_in_the_seed_round(chunk_buffers):
    """Seeds of chunks without a message start (sm_kernels.cuh: OOKD_SEED_IDLE / _NONE): long silences (tens of chunks
    without a transition) and messages spanning several chunks (a nexa message is 0.2 s = 5..10 chunks here) must come out
    as the reference's, without repair rounds (one host synchronisation, the seed round only)."""
    dev = O.load_device("p3l-nexa2012")
    stages = O.load_filter("fs32_fs4")
    msgs = [O.message_bytes(dev, util.nexa_fields(i)) for i in range(4)]
    tog, total = O.toggles_from_messages(dev, msgs, util.FS, 3000000)         # a second of silence in front
    tog = np.array(tog, dtype=np.uint64)
    gap = 4000000
    cut = len(tog) // 2 // 2 * 2
    tog = np.concatenate([tog[:cut], tog[cut:] + np.uint64(gap)])            # ... more between messages 2 and 3
    total += gap + 2000000                                                    # ... and behind
    i_on, q_on = O.on_level(0.9, 0.3)
    iq = O.synth(total, tog, i_on, q_on, O.noise_scale_for_sigma(0.01), 5)
    spb = 8192
    ref = O.rx(iq, stages, dev, samples_per_buffer=spb)
    assert len(ref["msgs"]) == 4
    g = B.Gpu(filter_stages=stages, sm=util.sm_spec(dev, stages), samples_per_buffer=spb, sm_chunk_buffers=chunk_buffers)
    for rep in range(2):
        got = g.decode(iq)
        assert got["msgs"] == ref["msgs"]
        assert np.array_equal(g.edges()[1], ref["edges"])
        assert got["sm_rounds"] == 1 and got["host_syncs"] == 1, (got["sm_rounds"], got["host_syncs"])
    g.close()


@pytest.mark.parametrize("seed", range(16))
def test_subwindow_fuzz(seed):
    """Seeded random shapes: device, filter, samples per buffer, chunk size, K, noise level, glitch bursts that make the
    reference lose messages (strings of dropped buffers across sub-window boundaries); messages, transitions and the exit
    state must be the oracle's / the undivided decode's."""
    rng = np.random.default_rng(1000 + seed)
    devname = ["p3l-nexa2012", "unknown-remote1"][int(rng.integers(2))]
    filt = ["fs32_fs4", "fs128_fs16_dec4", "fs64_fs8"][int(rng.integers(3))]
    spb = [1001, 4096, 8192][int(rng.integers(3))]
    k = int(rng.integers(2, 7))
    cb = int(rng.integers(2, 9))
    sigma = float(rng.choice([0.0, 0.02, 0.04, 0.06]))
    dev = O.load_device(devname)
    fields = util.nexa_fields if "nexa" in devname else util.remote_fields
    n_msgs = int(rng.integers(8, 14))
    iq0, _, tog = util.capture(dev, n_msgs, sigma=0.0, amplitude=0.8, phase=0.4, seed=seed, fields=fields)
    glitches = tuple((int(rng.integers(1000, len(iq0) - 2000)), int(rng.integers(20, 400))) for _ in range(int(rng.integers(0, 5))))
    iq, _, _ = util.capture(dev, n_msgs, sigma=sigma, amplitude=float(rng.uniform(0.5, 0.95)), phase=float(rng.uniform(0, 6.28)),
                            seed=seed, fields=fields, glitches=glitches)
    stages = O.load_filter(filt)
    sm = util.sm_spec(dev, stages)
    ref = O.rx(iq, stages, dev, samples_per_buffer=spb)
    plain = B.Gpu(filter_stages=stages, sm=sm, samples_per_buffer=spb, sm_chunk_buffers=cb)
    sub = B.Gpu(filter_stages=stages, sm=sm, samples_per_buffer=spb, sm_chunk_buffers=cb, sub_windows=k)
    want, wexit = plain.decode_shard(iq, 0, len(iq), True)
    assert want["msgs"] == ref["msgs"]
    for rep in range(2):
        got, gexit = sub.decode_shard(iq, 0, len(iq), True)
        assert got["msgs"] == ref["msgs"], (devname, filt, spb, k, cb, sigma, glitches)
        assert gexit == wexit
        assert np.array_equal(sub.edges()[1], ref["edges"])
    plain.close()
    sub.close()
