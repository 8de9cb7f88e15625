"""GPU parity: every result of the C ABI against the CPU oracle on identical inputs (bit-exact)."""
import numpy as np
import pytest

from oracle import oracle as O
from ookiedokie_b200 import binding as B
import ookd_testutil as util

pytestmark = pytest.mark.gpu

FILTERS = ["fs32_fs4", "fs128_fs16_dec4", "fs64_fs8", "unity1", "unity16", None]


def _gpu(filt, device=None, **kw):
    stages = O.load_filter(filt) if filt else None
    sm = util.sm_spec(device, stages) if device else None
    return B.Gpu(filter_stages=stages, sm=sm, **kw), stages


def test_synth_matches_oracle():
    dev = O.load_device("p3l-nexa2012")
    msgs = [O.message_bytes(dev, util.nexa_fields(i)) for i in range(2)]
    tog, total = O.toggles_from_messages(dev, msgs, util.FS, 12000)
    for scale, seed in [(0, 0), (O.noise_scale_for_sigma(0.02), 7), (O.noise_scale_for_sigma(0.3), 0xC0FFEE)]:
        ref = O.synth(total, tog, 1376, 1375, scale, seed)
        got = B.synth(total, tog, 1376, 1375, scale, seed)
        assert np.array_equal(ref, got)
    # twelve-term (near-Gaussian) noise
    sc = O.noise_scale_for_sigma(0.1, 12)
    ref = O.synth(total, tog, 600, -100, sc, 11, noise_terms=12)
    got = B.synth(total, tog, 600, -100, sc, 11, noise_terms=12)
    assert np.array_equal(ref, got)
    resid = ref[:12000, 0].astype(np.float64)                 # lead-in silence: pure noise
    assert abs(resid.std() / 204.8 - 1.0) < 0.05 and np.abs(resid).max() > 3.5 * 204.8
    # offset window
    ref = O.synth(5000, tog, 1945, 0, O.noise_scale_for_sigma(0.05), 3, first_sample=11000)
    got = B.synth(5000, tog, 1945, 0, O.noise_scale_for_sigma(0.05), 3, first_sample=11000)
    assert np.array_equal(ref, got)


@pytest.mark.parametrize("filt", FILTERS)
def test_filtered_samples_bit_exact(filt):
    rng = np.random.default_rng(5)
    iq = rng.integers(-2048, 2048, size=(70001, 2), dtype=np.int16)
    iq[:100] = 0
    iq[50] = (2047, -2048)          # impulse-ish start (gen_samples.m: impulse at sample 50)
    g, stages = _gpu(filt, samples_per_buffer=1024)
    got = g.filtered(iq)
    ref = O.rx(iq, stages, None, samples_per_buffer=1024, want_filtered=True)["filtered"]
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))     # bit-exact, stronger than 1e-5


@pytest.mark.parametrize("filt", ["fs32_fs4", "fs128_fs16_dec4", "unity16"])
def test_filter_cf_matches_fir_semantics(filt):
    rng = np.random.default_rng(11)
    x = rng.standard_normal((5000, 2)).astype(np.float32)
    g, stages = _gpu(filt)
    got = g.filter_cf(x)
    ref = O.Fir(stages).run(x)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


CASES = [
    # device, filter, n_msgs, sigma, amp, spb, flags
    ("p3l-nexa2012", "fs32_fs4", 3, 0.0, 0.95, 8192, 0),
    ("p3l-nexa2012", "fs32_fs4", 3, 0.02, 0.95, 8192, 0),
    ("p3l-nexa2012", "fs32_fs4", 3, 0.02, 0.95, 8192, B.FLAG_FORCE_GENERIC),
    ("p3l-nexa2012", "fs32_fs4", 4, 0.03, 0.95, 8192, 0),
    ("p3l-nexa2012", "fs32_fs4", 3, 0.02, 0.95, 1024, 0),
    ("p3l-nexa2012", "fs32_fs4", 3, 0.02, 0.95, 65536, 0),
    ("p3l-nexa2012", "fs32_fs4", 2, 0.02, 0.95, 5000, 0),
    ("p3l-nexa2012", "fs128_fs16_dec4", 3, 0.02, 0.95, 8192, 0),
    ("p3l-nexa2012", None, 2, 0.0, 0.95, 8192, 0),
    ("p3l-nexa2012", "fs64_fs8", 3, 0.05, 0.95, 8192, 0),
    ("unknown-remote1", "fs128_fs16_dec4", 10, 0.10, 0.30, 8192, 0),
    ("unknown-remote1", "fs128_fs16_dec4", 6, 0.05, 0.30, 8192, 0),
    ("unknown-remote1", "fs32_fs4", 4, 0.02, 0.5, 4096, 0),
    ("unknown-remote1", "fs128_fs16_dec4", 4, 0.05, 0.30, 1001, 0),     # spb not a multiple of the decimation
]


@pytest.mark.parametrize("devname,filt,n_msgs,sigma,amp,spb,flags", CASES)
def test_decode_matches_oracle(devname, filt, n_msgs, sigma, amp, spb, flags):
    dev = O.load_device(devname)
    fields = util.nexa_fields if "nexa" in devname else util.remote_fields
    iq, msgs, _ = util.capture(dev, n_msgs, sigma=sigma, amplitude=amp, phase=0.7, seed=n_msgs * 31 + spb, fields=fields)
    g, stages = _gpu(filt, dev, samples_per_buffer=spb, flags=flags, sm_chunk_buffers=7)
    ref = O.rx(iq, stages, dev, samples_per_buffer=spb, want_bits=True)
    got = g.decode(iq)
    assert got["n_out"] == ref["n_out"] and got["n_buffers"] == ref["n_buffers"]
    assert np.array_equal(g.bits(), ref["bits"])
    fb, edges = g.edges()
    assert fb == ref["first_bit"]
    assert np.array_equal(edges, ref["edges"])
    assert got["msgs"] == ref["msgs"]
    if sigma <= 0.02 and amp > 0.9:
        assert [m[3] for m in got["msgs"]] == msgs      # clean enough: every transmitted message decodes


def _piecewise_capture(rng, n, levels, sigma_lsb, seg=(200, 5000), full_range=False):
    """Carrier whose amplitude jumps between the given levels (fractions of full scale), random phase,
    rounded Gaussian noise.  Built with numpy only (no device model needed: decisions/edges are compared)."""
    amp = np.empty(n, np.float64)
    pos = 0
    while pos < n:
        ln = int(rng.integers(seg[0], seg[1]))
        amp[pos:pos + ln] = rng.choice(levels)
        pos += ln
    ph = rng.uniform(0, 2 * np.pi)
    i = amp * np.cos(ph) * 2048 + rng.normal(0, sigma_lsb, n)
    q = amp * np.sin(ph) * 2048 + rng.normal(0, sigma_lsb, n)
    lim = 32767 if full_range else 2047
    return np.stack([np.clip(np.rint(i), -lim - 1, lim), np.clip(np.rint(q), -lim - 1, lim)], 1).astype(np.int16)


SCREEN_CASES = [
    # levels, sigma (LSB), threshold, full_range
    ([0.0, 0.95], 41.0, 0.1, False),                 # the benchmark regime
    ([0.0, 0.95], 0.0, 0.1, False),                  # clean
    ([0.0, 0.1, 0.95], 5.0, 0.1, False),             # a level sitting on the threshold
    ([0.09, 0.1, 0.11, 0.0999, 0.1001], 1.0, 0.1, False),
    ([0.0, 0.3], 205.0, 0.1, False),                 # low SNR: almost every tile goes to the dense list
    ([0.0, 0.05, 0.2, 0.5], 20.0, 0.25, False),
    ([0.0, 4.0, 15.0], 300.0, 0.7, True),            # beyond the 12-bit range
    ([0.0], 0.0, 0.1, False),                        # all zero
    ([0.5], 0.0, 0.5, False),                        # constant exactly at the threshold
]


@pytest.mark.parametrize("filt", ["fs32_fs4", "fs64_fs8", "fs128_fs16_dec4"])
@pytest.mark.parametrize("levels,sigma,thr,full", SCREEN_CASES)
def test_screen_decisions_equal_exact(filt, levels, sigma, thr, full):
    rng = np.random.default_rng(abs(hash((filt, tuple(levels), sigma, thr))) % (2 ** 32))
    iq = _piecewise_capture(rng, 300001 if "dec4" in filt else 300000, levels, sigma, full_range=full)
    stages = O.load_filter(filt)
    ref = O.rx(iq, stages, None, threshold_=thr, samples_per_buffer=8192, want_bits=True)
    for flags in (0, B.FLAG_NO_TMA, B.FLAG_NO_SCREEN, B.FLAG_FORCE_GENERIC, B.FLAG_FMA_SCREEN):
        g = B.Gpu(filter_stages=stages, threshold=thr, samples_per_buffer=8192, flags=flags)
        got = g.decode(iq)
        assert np.array_equal(g.bits(), ref["bits"]), (flags, got["refined_blocks"], got["refined_tiles"])
        fb, edges = g.edges()
        assert fb == ref["first_bit"] and np.array_equal(edges, ref["edges"])


@pytest.mark.parametrize("filt,dec", [("fs32_fs4", 1), ("fs128_fs16_dec4", 4)])
def test_screen_actually_screens(filt, dec):
    # in the benchmark regime nearly every group of 8 outputs must be decided without MACs
    rng = np.random.default_rng(1)
    iq = _piecewise_capture(rng, 1 << 20, [0.0, 0.95], 41.0, seg=(1500, 12000))
    g = B.Gpu(filter_stages=O.load_filter(filt), threshold=0.1)
    got = g.decode(iq)
    groups = (1 << 20) // dec // 8
    assert got["refined_tiles"] == 0
    assert 0 < got["refined_blocks"] < 0.1 * groups, got


def test_screen_overflow_falls_back_to_exact_and_stays_correct():
    # everything near the threshold: the energy proofs decide nothing.  By default the probe kernel sees that and the
    # decode uses FMA screening from the start; without it (OOKD_FLAG_NO_ADAPTIVE) the work list overflows once and
    # the handle switches to FMA screening for good.  Decisions equal the oracle's either way.
    rng = np.random.default_rng(2)
    iq = _piecewise_capture(rng, 1 << 21, [0.1], 30.0)
    for filt in ("fs32_fs4", "fs128_fs16_dec4"):
        stages = O.load_filter(filt)
        ref = O.rx(iq, stages, None, threshold_=0.1, samples_per_buffer=8192, want_bits=True)
        g = B.Gpu(filter_stages=stages, threshold=0.1)
        first = g.decode(iq)
        assert first["fir_mode"] == 2 and first["refined_tiles"] == 0
        assert np.array_equal(g.bits(), ref["bits"])
        # levels sitting on the threshold: the rounding band really is visited, but by few groups
        assert 0 < first["refined_blocks"] < 0.05 * (1 << 21) / 8
        g = B.Gpu(filter_stages=stages, threshold=0.1, flags=B.FLAG_NO_ADAPTIVE)
        first = g.decode(iq)
        # the list holds 1/8 of all groups + 64 Ki: 2^21 fs32_fs4 outputs overflow it, 2^19 dec4 outputs do not
        assert first["refined_tiles"] == (1 if filt == "fs32_fs4" else 0)
        assert np.array_equal(g.bits(), ref["bits"])
        again = g.decode(iq)                      # after an overflow: FMA screening from the start
        assert again["refined_tiles"] == 0 and np.array_equal(g.bits(), ref["bits"])
        assert again["fir_mode"] == (2 if filt == "fs32_fs4" else 1)


@pytest.mark.parametrize("filt,log2n", [("fs32_fs4", 28), ("fs128_fs16_dec4", 27)])
def test_large_capture_paths_agree_and_are_deterministic(filt, log2n):
    """2^27..2^28 samples synthesised on the device: the persistent TMA-staged screening kernel (many tiles per CTA,
    where a shared-memory ring is reused), the FMA screen, the guarded-load path and the exact tiled kernel must give identical
    edge lists and messages, run after run (a ring reuse race only shows at this scale)."""
    import torch
    from ookiedokie_b200 import host as H
    n = 1 << log2n
    fir = H.Fir(filt)
    dev = H.Device("p3l-nexa2012", util.FS // fir.total_decimation)
    txdev = H.Device("p3l-nexa2012", util.FS)                 # the transmitter runs at the full sample rate
    msgs = [txdev.message({"Channel": str(1 + i % 3), "Temperature (C)": f"{-20.0 + 0.1 * ((i * 37) % 900):.1f}"})
            for i in range(n // 380000 + 8)]
    tog, total = txdev.toggles(msgs, 12000)
    assert total >= n
    d_iq = torch.empty((n * 2,), dtype=torch.int16, device="cuda")
    B.synth(n, np.ascontiguousarray(tog), 1488, 1253, O.noise_scale_for_sigma(0.02), 0xC0FFEE, device_id=0,
            device_ptr=d_iq.data_ptr())
    torch.cuda.synchronize()
    ref = None
    for flags, reps in [(B.FLAG_NO_SCREEN, 1), (B.FLAG_FMA_SCREEN, 1), (B.FLAG_NO_TMA, 1), (0, 4)]:
        g = B.Gpu(filter_stages=fir.stages, sm=dev.sm_spec(), threshold=0.1, samples_per_buffer=8192, flags=flags)
        g.want_list = False
        for _ in range(reps):
            res, _ = g.decode_shard((d_iq.data_ptr(), n), 0, n, True, None)
            fb, edges = g.edges()
            cur = (fb, edges.copy(), res["msgs_raw"].copy())
            if ref is None:
                ref = cur
                assert len(cur[2]) > 200
            assert cur[0] == ref[0] and np.array_equal(cur[1], ref[1]) and np.array_equal(cur[2], ref[2]), flags
        g.close()


def test_batch_of_independent_captures_matches_oracle():
    """BASELINE configs[4] in miniature: mixed devices through fs64_fs8, several captures in flight on several
    handles (ookd_gpu_batch_decode); every capture must decode exactly like the oracle run on it alone."""
    stages = O.load_filter("fs64_fs8")
    devs = [O.load_device("p3l-nexa2012"), O.load_device("unknown-remote1")]
    gpus = []
    for d in devs:
        for _ in range(2):                       # two handles per device description: two captures of a kind in flight
            gpus.append(B.Gpu(filter_stages=stages, sm=util.sm_spec(d, stages), threshold=0.1, samples_per_buffer=8192))
    captures, want = [], []
    for i in range(10):
        kind = i % 2
        dev = devs[kind]
        sigma = [0.0, 0.02, 0.05][i % 3]
        fields = util.nexa_fields if kind == 0 else None
        iq, sent, _ = util.capture(dev, 2 + i % 3, sigma=sigma, phase=0.3 * i, seed=100 + i, fields=fields)
        captures.append((iq, 2 * kind + (i // 2) % 2))
        want.append(O.rx(iq, stages, dev, samples_per_buffer=8192)["msgs"])
    msgs, stats = B.batch_decode(gpus, captures)
    for i in range(10):
        nbytes = (devs[i % 2]["num_bits"] + 7) // 8
        assert B.msgs_to_tuples(msgs[i], nbytes) == [tuple(m) for m in want[i]], i
        assert stats[i]["gpu_launches"] > 0
    assert sum(len(m) for m in msgs) > 0
    # a second batch on the same handles, device-resident inputs this time
    import torch
    d_caps = [torch.from_numpy(np.ascontiguousarray(c[0]).reshape(-1)).cuda() for c in captures[:4]]
    msgs2, _ = B.batch_decode(gpus, [((t.data_ptr(), t.numel() // 2), captures[i][1]) for i, t in enumerate(d_caps)])
    for i in range(4):
        assert np.array_equal(msgs2[i], msgs[i])


@pytest.mark.parametrize("devname,filt", [("p3l-nexa2012", "fs32_fs4"), ("unknown-remote1", "fs128_fs16_dec4")])
def test_every_code_path_gives_the_same_decode(devname, filt):
    """Flags select alternative kernels / host paths (tensor-copy vs guarded-load screening, FMA screening, exact
    kernel, synchronous tail, shared SMs, generic FIR): edges and messages must be identical, and equal to the oracle."""
    dev = O.load_device(devname)
    fields = util.nexa_fields if "nexa" in devname else util.remote_fields
    iq, sent, _ = util.capture(dev, 12, sigma=0.02, amplitude=0.6, phase=1.1, seed=4242, fields=fields,
                               glitches=((9000, 100),))
    stages = O.load_filter(filt)
    ref = O.rx(iq, stages, dev, samples_per_buffer=8192, want_bits=True)
    for flags in (0, B.FLAG_SYNC_TAIL, B.FLAG_NO_TMA, B.FLAG_NO_SCREEN, B.FLAG_SHARE_SMS, B.FLAG_FMA_SCREEN, B.FLAG_NO_ADAPTIVE,
                  B.FLAG_FORCE_GENERIC, B.FLAG_NO_TMA | B.FLAG_SYNC_TAIL, B.FLAG_NO_GRAPH, B.FLAG_FUSED_SM,
                  B.FLAG_FUSED_SM | B.FLAG_NO_GRAPH):
        for chunk_buffers in (0, 5):
            g = B.Gpu(filter_stages=stages, sm=util.sm_spec(dev, stages), threshold=0.1, samples_per_buffer=8192, flags=flags,
                      sm_chunk_buffers=chunk_buffers)
            got = g.decode(iq)
            fb, edges = g.edges()
            assert fb == ref["first_bit"] and np.array_equal(edges, ref["edges"]), (flags, chunk_buffers)
            assert got["msgs"] == ref["msgs"], (flags, chunk_buffers)
            g.close()
    assert len(ref["msgs"]) >= 8


def test_two_decodes_in_flight():
    """ookd_gpu_decode_begin on two handles, then ookd_gpu_decode_end on both (and the other way round)."""
    dev = O.load_device("p3l-nexa2012")
    stages = O.load_filter("fs32_fs4")
    caps = [util.capture(dev, 4 + i, sigma=0.02, phase=0.5 * i, seed=900 + i, fields=util.nexa_fields)[0] for i in range(2)]
    want = [O.rx(c, stages, dev, samples_per_buffer=8192)["msgs"] for c in caps]
    gpus = [B.Gpu(filter_stages=stages, sm=util.sm_spec(dev, stages), threshold=0.1, samples_per_buffer=8192) for _ in range(2)]
    for order in ((0, 1), (1, 0)):
        for i in range(2):
            gpus[i].decode_begin(caps[i], 0, len(caps[i]), True)
        with pytest.raises(B.OokdError):
            gpus[0].decode_begin(caps[0], 0, len(caps[0]), True)       # one decode per handle
        for i in order:
            res, _ = gpus[i].decode_end()
            assert res["msgs"] == want[i]
    with pytest.raises(B.OokdError):
        gpus[0].decode_end()                                           # nothing in flight


def test_graph_replay_across_different_captures_of_one_shape():
    """The decode tail is captured into a CUDA graph once per geometry and replayed: same-length captures with
    different contents (and a different length in between, which re-captures) must each decode like the oracle."""
    dev = O.load_device("p3l-nexa2012")
    stages = O.load_filter("fs32_fs4")
    g = B.Gpu(filter_stages=stages, sm=util.sm_spec(dev, stages), threshold=0.1, samples_per_buffer=8192)
    base = [util.capture(dev, 5, sigma=0.02, phase=0.2 * i, seed=300 + i, fields=util.nexa_fields)[0] for i in range(4)]
    n = min(len(c) for c in base)
    caps = [c[:n] for c in base] + [base[0][: n - 50000]] + [c[:n] for c in base[:2]]
    for c in caps:
        want = O.rx(c, stages, dev, samples_per_buffer=8192)
        got = g.decode(c)
        assert got["msgs"] == want["msgs"]
        assert np.array_equal(g.edges()[1], want["edges"])
