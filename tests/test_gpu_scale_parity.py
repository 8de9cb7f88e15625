"""GPU parity AT SCALE against the CPU oracle (not GPU-vs-GPU): BASELINE configs[1] and configs[2] at sizes where the
persistent kernels, the chunk tables and the stitcher are really exercised, with true Gaussian noise besides the
benchmark's integer recipe, plus a seeded fuzz of the screening proofs.

Reference semantics being reproduced: threshold src/ookiedokie.c:171-179, device_process (drop the rest of the buffer
on ERROR) src/device.c:634-658, in-order MACs src/fir.c:311-318.
"""
import numpy as np
import pytest

from oracle import oracle as O
from ookiedokie_b200 import binding as B
import ookd_testutil as util

pytestmark = pytest.mark.gpu


def _envelope(tog, s0, s1):
    """level of samples [s0, s1): a toggle at position p changes the level from sample p on"""
    before = int(np.searchsorted(tog, np.uint64(s0), side="left"))
    loc = tog[(tog >= s0) & (tog < s1)].astype(np.int64) - s0
    d = np.zeros(s1 - s0, np.int64)
    np.add.at(d, loc, 1)
    return ((np.cumsum(d) + before) & 1).astype(np.float32)


def gaussian_capture(dev, n_msgs, fields, sigma, amp, phase, seed, samplerate=util.FS, lead=12000, tail=20000,
                     cfo_cycles_per_sample=0.0, dc=(0.0, 0.0)):
    """Device messages tiled back to back, carrier amp*e^{j(phase + 2 pi f n)}, AWGN (true Gaussian, sigma per
    component in full-scale units), DC offset, round to nearest, clip to the 12-bit SC16Q11 range."""
    msgs = [O.message_bytes(dev, fields(i) if fields else {}) for i in range(n_msgs)]
    tog, total = O.toggles_from_messages(dev, msgs, samplerate, lead)
    total += tail
    rng = np.random.default_rng(seed)
    out = np.empty((total, 2), np.int16)
    step = 1 << 24
    for s0 in range(0, total, step):
        s1 = min(total, s0 + step)
        env = _envelope(tog, s0, s1)
        ph = phase + 2.0 * np.pi * cfo_cycles_per_sample * np.arange(s0, s1, dtype=np.float64)
        i = env * (amp * 2048.0 * np.cos(ph)).astype(np.float32)
        q = env * (amp * 2048.0 * np.sin(ph)).astype(np.float32)
        i += rng.standard_normal(s1 - s0, dtype=np.float32) * np.float32(sigma * 2048.0) + np.float32(dc[0] * 2048.0)
        q += rng.standard_normal(s1 - s0, dtype=np.float32) * np.float32(sigma * 2048.0) + np.float32(dc[1] * 2048.0)
        out[s0:s1, 0] = np.clip(np.rint(i), -2048, 2047).astype(np.int16)
        out[s0:s1, 1] = np.clip(np.rint(q), -2048, 2047).astype(np.int16)
    return out, msgs


def _compare(g, got, ref, what):
    fb, edges = g.edges()
    assert fb == ref["first_bit"], what
    assert len(edges) == len(ref["edges"]) and np.array_equal(edges, ref["edges"]), what
    assert got["msgs"] == ref["msgs"], what


def test_c2_benchmark_recipe_2pow27_vs_oracle():
    """BASELINE configs[1] on a 2^27-sample prefix (512 MiB, ~330 messages): the bytes bench.py decodes (device-side
    integer synthesis), messages and edge list against the oracle."""
    import torch
    from ookiedokie_b200 import host as H
    n = 1 << 27
    fir = H.Fir("fs32_fs4")
    hdev = H.Device("p3l-nexa2012", util.FS)
    msgs = [hdev.message({"Channel": str(1 + i % 3), "Temperature (C)": f"{-20.0 + 0.1 * ((i * 37) % 900):.1f}"})
            for i in range(n // 380000 + 8)]
    tog, total = hdev.toggles(msgs, 12000)
    d_iq = torch.empty((n * 2,), dtype=torch.int16, device="cuda")
    B.synth(n, np.ascontiguousarray(tog), 1488, 1253, O.noise_scale_for_sigma(0.02), 0x00C0FFEE, device_id=0,
            device_ptr=d_iq.data_ptr())
    torch.cuda.synchronize()
    iq = d_iq.cpu().numpy().reshape(-1, 2)
    odev = O.load_device("p3l-nexa2012")
    stages = O.load_filter("fs32_fs4")
    ref = O.rx(iq, stages, odev, samples_per_buffer=8192)
    assert len(ref["msgs"]) > 300
    g = B.Gpu(filter_stages=stages, sm=util.sm_spec(odev, stages), threshold=0.1, samples_per_buffer=8192)
    for rep in range(2):                                           # device input, then host input
        got = g.decode((d_iq.data_ptr(), n)) if rep == 0 else g.decode(iq)
        _compare(g, got, ref, rep)


@pytest.mark.parametrize("sigma", [0.02, 0.03])
def test_c2_gaussian_noise_vs_oracle(sigma):
    """Same device/filter with TRUE Gaussian noise (tails beyond the +-3.46 sigma of the integer recipe produce
    glitch pulses that trigger the buffer-drop rule); sigma 0.03 is SURVEY 8(d)'s second parity case."""
    odev = O.load_device("p3l-nexa2012")
    stages = O.load_filter("fs32_fs4")
    iq, sent = gaussian_capture(odev, 150, util.nexa_fields, sigma, 0.95, 0.7, seed=int(sigma * 1000))
    ref = O.rx(iq, stages, odev, samples_per_buffer=8192)
    g = B.Gpu(filter_stages=stages, sm=util.sm_spec(odev, stages), threshold=0.1, samples_per_buffer=8192)
    got = g.decode(iq)
    _compare(g, got, ref, sigma)
    assert 0 < len(ref["msgs"]) <= len(sent)
    print(f"sigma {sigma}: {len(ref['msgs'])}/{len(sent)} messages, {len(ref['edges'])} edges, "
          f"refined {got['refined_blocks']} groups of {got['n_out'] // 8}")


def test_c3_lowsnr_1000_messages_vs_oracle():
    """BASELINE configs[2]: unknown-remote1 through fs128_fs16_dec4, amp 0.30, Gaussian sigma 0.10, >= 1000 messages
    transmitted; partial decode expected (the dropped-buffer cascades of device_process are what this stresses)."""
    odev = O.load_device("unknown-remote1")
    stages = O.load_filter("fs128_fs16_dec4")
    iq, sent = gaussian_capture(odev, 1050, util.remote_fields, 0.10, 0.30, 0.4, seed=7)
    ref = O.rx(iq, stages, odev, samples_per_buffer=8192)
    assert 100 < len(ref["msgs"]) < len(sent)                      # partial decode regime
    g = B.Gpu(filter_stages=stages, sm=util.sm_spec(odev, stages), threshold=0.1, samples_per_buffer=8192)
    for rep in range(2):                                           # (second decode: after the screen's verdict on the first)
        got = g.decode(iq)
        _compare(g, got, ref, rep)
    assert got["fir_mode"] == 2                                    # the probe kernel chose FMA screening by itself
    print(f"C3: {len(ref['msgs'])}/{len(sent)} messages, {len(ref['edges'])} edges, sm_rounds {got['sm_rounds']}, "
          f"refined_tiles {got['refined_tiles']}, refined groups {got['refined_blocks']}")


def test_many_messages_three_decodes_one_handle_and_slot_overflow():
    """> 1024 messages per decode, three decodes on one handle (pinned message staging regrown, tail graph
    re-captured), then chunks long enough to overflow the per-chunk message slots."""
    fs = 1000000
    odev = O.load_device("unknown-remote1")
    stages = O.load_filter("fs32_fs4")
    iq, sent, _ = util.capture(odev, 1300, sigma=0.02, amplitude=0.5, phase=0.3, seed=5, fields=util.remote_fields,
                               samplerate=fs, lead=4000)
    ref = O.rx(iq, stages, odev, samples_per_buffer=8192, samplerate=fs)
    assert len(ref["msgs"]) > 1024
    sm = dict(states=odev["states"], num_bits=odev["num_bits"], sample_rate=fs)
    g = B.Gpu(filter_stages=stages, sm=sm, threshold=0.1, samples_per_buffer=8192)
    for rep in range(3):
        got = g.decode(iq)
        _compare(g, got, ref, rep)
    # a shorter capture in between (different geometry), then the long one again
    cut = (len(iq) // 3) // 8192 * 8192
    ref_cut = O.rx(iq[:cut], stages, odev, samples_per_buffer=8192, samplerate=fs)
    _compare(g, g.decode(iq[:cut]), ref_cut, "cut")
    _compare(g, g.decode(iq), ref, "again")
    g.close()
    g2 = B.Gpu(filter_stages=stages, sm=sm, threshold=0.1, samples_per_buffer=8192, sm_chunk_buffers=512)
    for rep in range(2):
        _compare(g2, g2.decode(iq), ref, f"slots {rep}")


FUZZ_FILTERS = ["fs32_fs4", "fs64_fs8", "fs128_fs16_dec4"]


@pytest.mark.parametrize("filt", FUZZ_FILTERS)
def test_screen_fuzz_decisions_equal_oracle(filt):
    """Seeded fuzz of the screening proofs (>= 200 random captures per filter shape): Gaussian noise, carrier
    frequency offset up to +-Fs/64, DC offset, amplitudes hovering around the threshold, clipping at full scale.
    The decisions must equal the oracle's bit for bit whatever the screen proves or hands to the exact kernel."""
    stages = O.load_filter(filt)
    rng = np.random.default_rng(20261018 + FUZZ_FILTERS.index(filt))
    thrs = [0.1, 0.05, 0.25, 0.6]
    gpus = {t: B.Gpu(filter_stages=stages, threshold=t, samples_per_buffer=8192) for t in thrs}
    fma = {t: B.Gpu(filter_stages=stages, threshold=t, samples_per_buffer=8192, flags=B.FLAG_FMA_SCREEN) for t in thrs}
    n_cases, n = 208, 49152
    refined = refined_fma = 0
    for case in range(n_cases):
        thr = thrs[case % len(thrs)]
        kind = case % 8
        # amplitude profile: segments of random length; levels drawn around the threshold for some kinds
        amp = np.empty(n, np.float64)
        pos = 0
        while pos < n:
            ln = int(rng.integers(40, 6000))
            if kind in (0, 1):
                lvl = rng.choice([0.0, rng.uniform(0.5, 1.0)])
            elif kind in (2, 3):
                lvl = thr * rng.uniform(0.7, 1.4)                      # hovering around the threshold
            elif kind == 4:
                lvl = rng.choice([0.0, thr * rng.uniform(0.95, 1.05), 1.4])       # incl. beyond full scale (clips)
            else:
                lvl = rng.uniform(0.0, 1.2)
            amp[pos:pos + ln] = lvl
            pos += ln
        sigma = [0.0, 0.005, 0.02, 0.05, 0.1, 0.3][int(rng.integers(0, 6))]
        cfo = rng.uniform(-1.0 / 64, 1.0 / 64) if kind != 0 else 0.0
        dc = (rng.uniform(-0.05, 0.05), rng.uniform(-0.05, 0.05)) if kind in (5, 6) else (0.0, 0.0)
        ph = rng.uniform(0, 2 * np.pi) + 2 * np.pi * cfo * np.arange(n)
        i = amp * np.cos(ph) * 2048 + rng.normal(0, 1, n) * sigma * 2048 + dc[0] * 2048
        q = amp * np.sin(ph) * 2048 + rng.normal(0, 1, n) * sigma * 2048 + dc[1] * 2048
        lim = 32767 if kind == 7 else 2047                             # kind 7: the full int16 range
        if kind == 7:
            i *= 6.0
            q *= 6.0
        iq = np.stack([np.clip(np.rint(i), -lim - 1, lim), np.clip(np.rint(q), -lim - 1, lim)], 1).astype(np.int16)
        ref = O.rx(iq, stages, None, threshold_=thr, samples_per_buffer=8192, want_bits=True)
        g = gpus[thr]
        got = g.decode(iq)
        bits = g.bits()
        assert np.array_equal(bits, ref["bits"]), (filt, case, kind, thr, sigma, cfo, dc,
                                                    int(np.flatnonzero(bits != ref["bits"])[0]))
        fb, edges = g.edges()
        assert fb == ref["first_bit"] and np.array_equal(edges, ref["edges"]), (filt, case)
        refined += got["refined_blocks"]
        # FMA screening: fused multiply-add decisions + exact recomputation inside the rounding band
        gf = fma[thr]
        got_f = gf.decode(iq)
        assert got_f["fir_mode"] == 2
        bits_f = gf.bits()
        assert np.array_equal(bits_f, ref["bits"]), ("fma", filt, case, kind, thr, sigma,
                                                     int(np.flatnonzero(bits_f != ref["bits"])[0]))
        refined_fma += got_f["refined_blocks"]
    assert refined > 0
    print(f"{filt}: energy screen refined {refined} groups, fma screen {refined_fma} of {n_cases * n // 8} (per decimation)")
