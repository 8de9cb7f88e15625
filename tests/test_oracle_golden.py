"""Pins the CPU oracle (oracle/ookd_oracle.c) against vectors produced by the unmodified reference
(tools/make_golden.py -> tests/golden).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
import ookd_testutil as util

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILTERS = ["fs32_fs4", "fs128_fs16_dec4", "fs64_fs8", "unity1", "unity16"]
RX = json.load(open(os.path.join(GOLD, "rx.json")))
TX = json.load(open(os.path.join(GOLD, "tx.json")))


@pytest.mark.parametrize("filt", FILTERS)
def test_fir_matches_fir_test(filt):
    g = np.load(os.path.join(GOLD, f"fir_{filt}.npz"))
    stages = O.load_filter(filt)
    for key in g.files:
        if not key.startswith("in_"):
            continue
        want = g["out_" + key[3:]]
        for chunk in (len(g[key]), 7, 32):           # streaming state: result independent of chunking
            fir = O.Fir(stages)
            x = g[key]
            got = np.concatenate([fir.run(x[i:i + chunk]) for i in range(0, len(x), chunk)] or [np.zeros((0, 2), np.float32)])
            assert got.shape == want.shape, (filt, key)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (filt, key, chunk)


def test_impulse_response_is_taps():
    # fir_test property the reference's own stimuli pin: impulse response == taps as float32
    g = np.load(os.path.join(GOLD, "fir_fs32_fs4.npz"))
    taps = O.load_filter("fs32_fs4")[0][1]
    assert np.array_equal(g["out_imp_i"][49:49 + 32, 0], taps)
    assert np.array_equal(g["out_imp_q"][49:49 + 32, 1], taps)


@pytest.mark.parametrize("case", TX, ids=lambda c: f"{c['device']}-{c['count']}-{len(c['params'])}")
def test_tx_generator_matches_reference(case):
    dev = O.load_device(case["device"])
    data = O.message_bytes(dev, case["params"])
    lead = int(3000000 * case["delay_us"] // 1000000)
    tog, total = O.toggles_from_messages(dev, [data] * case["count"], 3000000, lead)
    iq = O.synth(total, tog, 1945, 0, 0, 0)
    assert total == case["n_samples"]
    level = (iq[:, 0] != 0).astype(np.int8)
    change = np.flatnonzero(np.diff(level)) + 1
    starts = np.concatenate([[0], change])
    lens = np.diff(np.concatenate([starts, [len(level)]]))
    runs = [[int(iq[s, 0]), int(iq[s, 1]), int(n)] for s, n in zip(starts, lens)]
    assert runs == case["runs"]


def build_capture(case):
    dev = O.load_device(case["device"])
    fields = util.nexa_fields if "nexa" in case["device"] else util.remote_fields
    iq, msgs, _ = util.capture(dev, case["n_msgs"], sigma=case["sigma"], amplitude=case["amp"], phase=case["phase"],
                               seed=case["seed"], fields=fields, glitches=[tuple(g) for g in case.get("glitches", [])])
    return dev, iq, msgs


def filter_of(case):
    if case["filter"] is None:
        return O.load_filter("fs128_fs16_dec4")     # sdr_default_filter for bladerf_file
    if case["filter"] == "none":
        return None
    return O.load_filter(case["filter"])


@pytest.mark.parametrize("case", RX, ids=lambda c: c["name"])
def test_rx_matches_reference(case):
    from ookiedokie_b200 import host as H
    dev, iq, msgs = build_capture(case)
    assert len(iq) == case["n_samples"] and [m.hex() for m in msgs] == case["sent"]
    stages = filter_of(case)
    res = O.rx(iq, stages, dev, threshold_=case["thr"], samples_per_buffer=case["spb"])
    assert res["first_bit"] == case["first_bit"]
    assert [int(e) for e in res["edges"]] == case["edges"]
    # decoded rows: format the oracle's message bytes with the host formatter, group per buffer
    dec = O.filter_total_decimation(stages) if stages else 1
    hdev = H.Device(case["device"], 3000000 // dec)
    rows, header, cur_buf, cur = [], None, None, None
    for out_sample, buf, nbits, data in res["msgs"]:
        kv = [(k, v) for k, v in hdev.format(data) if k != "Decode Timestamp"]
        if buf != cur_buf:
            cur = []
            rows.append(cur)
            cur_buf = buf
        if header is None:
            header = [k for k, _ in kv]
        cur.extend(v for _, v in kv)
    want = case["csv_rows"]
    if not want:
        assert rows == []
    else:
        assert header == want[0]
        assert rows == want[1:]


REC = json.load(open(os.path.join(GOLD, "rec.json")))


def recorder_bytes(case, rec_input):
    """What the reference's recorder writes (src/ookiedokie.c:248-270 + complexf_to_sc16q11, src/complexf.h:87-96),
    restated on the oracle's filtered samples."""
    dev, iq, _ = build_capture(case)
    stages = None if case["filter"] == "none" else O.load_filter(case["filter"] or "fs128_fs16_dec4")
    spb = case["spb"]
    if rec_input or stages is None:                     # no filter forces input recording (src/main.c:668-671)
        n_eff = (len(iq) + spb - 1) // spb * spb
        out = np.zeros((n_eff, 2), np.int16)
        out[:len(iq)] = iq                              # int16 -> float -> (int16_t)(x * 2048.0f) is the identity
        return out
    f = O.rx(iq, stages, None, samples_per_buffer=spb, want_filtered=True)["filtered"]
    return np.trunc(f * np.float32(2048.0)).astype(np.int32).astype(np.int16)


@pytest.mark.parametrize("rec", REC, ids=lambda r: r["name"])
def test_recorder_restatement_matches_reference(rec):
    import hashlib
    case = next(c for c in RX if c["name"] == rec["name"])
    out = recorder_bytes(case, rec["rec_input"])
    assert len(out) == rec["n_samples"]
    assert out[rec["probe_at"]:rec["probe_at"] + 16].reshape(-1).tolist() == rec["probe"]
    assert hashlib.sha256(out.tobytes()).hexdigest() == rec["sha256"]
