#!/usr/bin/env python3
"""Generate tests/golden/* by running the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile).  Runs in the build container only; the outputs are committed
so the GPU box (no /root/reference) can check both the oracle and the CUDA path against them.

What is pinned:
  fir_*.npz   fir_test (reference src/test/fir_test.c) outputs for the shipped filter shapes on the
              stimuli of src/matlab/gen_samples.m (impulses, Fs/4 and Fs/32 tones, two-tone) plus
              noise, identical across chunk sizes 32 / 7 / 8192
  tx.json     `ookiedokie --tx bladerf_file` captures as run lengths
  rx.json     `ookiedokie --rx bladerf_file ... --rx-fmt csv -B edges` on synthetic captures
              (clean, noisy, low-SNR, glitch / samples-per-buffer dependence): threshold edges and
              decoded rows (the wall-clock "Decode Timestamp" column dropped)
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
import ookd_testutil as util  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF = O.ref_binary("ookiedokie")
FIR_TEST = O.ref_binary("fir_test")


def fir_stimuli():
    rng = np.random.default_rng(2015)
    n = 100
    imp_i = np.zeros((n, 2), np.float32); imp_i[49, 0] = 1.0        # MATLAB index 50
    imp_q = np.zeros((n, 2), np.float32); imp_q[49, 1] = 1.0
    t = np.arange(4000)
    tone4 = np.stack([np.cos(2 * np.pi * t / 4), np.sin(2 * np.pi * t / 4)], 1).astype(np.float32)
    tone32 = np.stack([np.cos(2 * np.pi * t / 32), np.sin(2 * np.pi * t / 32)], 1).astype(np.float32)
    two = (0.5 * (tone4 + tone32)).astype(np.float32)
    long_imp = np.zeros((20000, 2), np.float32); long_imp[999, 0] = 1.0
    noise = (rng.integers(-2048, 2048, size=(5000, 2)).astype(np.float32) * np.float32(1.0 / 2048.0))
    return dict(imp_i=imp_i, imp_q=imp_q, tone4=tone4, tone32=tone32, two=two, long_imp=long_imp, noise=noise)


def make_fir():
    stim = fir_stimuli()
    for filt in ["fs32_fs4", "fs128_fs16_dec4", "fs64_fs8", "unity1", "unity16"]:
        out = {}
        for name, x in stim.items():
            res = []
            for chunk in (32, 7, 8192):
                with tempfile.TemporaryDirectory() as td:
                    fin, fout = os.path.join(td, "in.cf"), os.path.join(td, "out.cf")
                    x.tofile(fin)
                    subprocess.run([FIR_TEST, filt, fin, fout, str(chunk)], check=True, capture_output=True)
                    res.append(np.fromfile(fout, dtype=np.float32).reshape(-1, 2))
            assert all(np.array_equal(res[0].view(np.uint32), r.view(np.uint32)) for r in res[1:]), (filt, name)
            out["in_" + name] = x
            out["out_" + name] = res[0]
        np.savez_compressed(os.path.join(GOLD, f"fir_{filt}.npz"), **out)
        print("fir", filt, {k: v.shape for k, v in out.items() if k.startswith("out_")})


def runs_of(iq):
    level = (iq[:, 0] != 0).astype(np.int8)
    change = np.flatnonzero(np.diff(level)) + 1
    starts = np.concatenate([[0], change])
    lens = np.diff(np.concatenate([starts, [len(level)]]))
    return [[int(iq[s, 0]), int(iq[s, 1]), int(n)] for s, n in zip(starts, lens)]


def make_tx():
    cases = [("p3l-nexa2012", {}, 1, 4000), ("p3l-nexa2012", {"Channel": "2", "Temperature (C)": "21.5"}, 3, 4000),
             ("p3l-nexa2012", {"Channel": "3", "Temperature (C)": "-12.3", "Unknown-2": "0xa5"}, 2, 1000),
             ("unknown-remote1", {}, 1, 4000), ("unknown-remote1", {"ID": "0x42", "Button": "Pause"}, 3, 4000),
             ("unknown-remote1", {"Button": "0x1234"}, 1, 0)]
    out = []
    for dev, params, count, delay in cases:
        with tempfile.TemporaryDirectory() as td:
            cap = os.path.join(td, "c.sc16q11")
            args = [REF, "--tx", "bladerf_file", "-A", cap, "-d", dev, "-c", str(count), "-D", str(delay)]
            for k, v in params.items():
                args += ["-p", f"{k}={v}"]
            subprocess.run(args, check=True, capture_output=True)
            iq = np.fromfile(cap, dtype=np.int16).reshape(-1, 2)
        out.append(dict(device=dev, params=params, count=count, delay_us=delay, n_samples=len(iq), runs=runs_of(iq)))
        print("tx", dev, params, len(iq))
    json.dump(out, open(os.path.join(GOLD, "tx.json"), "w"), indent=0)


RX_CASES = [
    # name, device, filter(None => default, "none" => off), n_msgs, sigma, amp, phase, seed, spb, thr, glitches
    dict(name="nexa_clean_fs32", device="p3l-nexa2012", filter="fs32_fs4", n_msgs=3, sigma=0.0, amp=0.95, phase=0.0, seed=1, spb=8192, thr=0.1),
    dict(name="nexa_awgn02_fs32", device="p3l-nexa2012", filter="fs32_fs4", n_msgs=4, sigma=0.02, amp=0.95, phase=0.7, seed=2, spb=8192, thr=0.1),
    dict(name="nexa_awgn03_fs32", device="p3l-nexa2012", filter="fs32_fs4", n_msgs=6, sigma=0.03, amp=0.95, phase=2.1, seed=3, spb=8192, thr=0.1),
    dict(name="nexa_default_filter", device="p3l-nexa2012", filter=None, n_msgs=3, sigma=0.02, amp=0.95, phase=-1.0, seed=4, spb=8192, thr=0.1),
    dict(name="nexa_nofilter", device="p3l-nexa2012", filter="none", n_msgs=2, sigma=0.0, amp=0.95, phase=0.3, seed=5, spb=8192, thr=0.1),
    dict(name="nexa_fs64", device="p3l-nexa2012", filter="fs64_fs8", n_msgs=3, sigma=0.05, amp=0.95, phase=0.0, seed=6, spb=8192, thr=0.1),
    dict(name="nexa_spb1024", device="p3l-nexa2012", filter="fs32_fs4", n_msgs=3, sigma=0.02, amp=0.95, phase=0.7, seed=7, spb=1024, thr=0.1),
    dict(name="nexa_spb65536", device="p3l-nexa2012", filter="fs32_fs4", n_msgs=3, sigma=0.02, amp=0.95, phase=0.7, seed=8, spb=65536, thr=0.1),
    dict(name="nexa_spb5000", device="p3l-nexa2012", filter="fs128_fs16_dec4", n_msgs=2, sigma=0.02, amp=0.95, phase=0.7, seed=9, spb=5000, thr=0.1),
    dict(name="nexa_thr03", device="p3l-nexa2012", filter="fs32_fs4", n_msgs=3, sigma=0.05, amp=0.95, phase=0.7, seed=10, spb=8192, thr=0.3),
    dict(name="nexa_glitch_spb8192", device="p3l-nexa2012", filter="fs32_fs4", n_msgs=2, sigma=0.0, amp=0.95, phase=0.0, seed=11, spb=8192, thr=0.1, glitches=[[9000, 100]]),
    dict(name="nexa_glitch_spb1024", device="p3l-nexa2012", filter="fs32_fs4", n_msgs=2, sigma=0.0, amp=0.95, phase=0.0, seed=11, spb=1024, thr=0.1, glitches=[[9000, 100]]),
    dict(name="remote1_lowsnr_dec4", device="unknown-remote1", filter=None, n_msgs=12, sigma=0.10, amp=0.30, phase=0.9, seed=12, spb=8192, thr=0.1),
    dict(name="remote1_snr_dec4", device="unknown-remote1", filter="fs128_fs16_dec4", n_msgs=8, sigma=0.05, amp=0.30, phase=0.9, seed=13, spb=8192, thr=0.1),
    dict(name="remote1_fs32", device="unknown-remote1", filter="fs32_fs4", n_msgs=4, sigma=0.02, amp=0.5, phase=0.0, seed=14, spb=4096, thr=0.1),
    dict(name="remote1_spb1001", device="unknown-remote1", filter="fs128_fs16_dec4", n_msgs=4, sigma=0.05, amp=0.30, phase=0.0, seed=15, spb=1001, thr=0.1),
]


def build_capture(case):
    dev = O.load_device(case["device"])
    fields = util.nexa_fields if "nexa" in case["device"] else util.remote_fields
    iq, msgs, tog = util.capture(dev, case["n_msgs"], sigma=case["sigma"], amplitude=case["amp"], phase=case["phase"],
                                 seed=case["seed"], fields=fields, glitches=[tuple(g) for g in case.get("glitches", [])])
    return dev, iq, msgs


def make_rx():
    out = []
    for case in RX_CASES:
        dev, iq, msgs = build_capture(case)
        with tempfile.TemporaryDirectory() as td:
            cap, dig = os.path.join(td, "c.sc16q11"), os.path.join(td, "dig.csv")
            iq.tofile(cap)
            args = [REF, "--rx", "bladerf_file", "-A", cap, "-d", case["device"], "--rx-fmt", "csv", "-B", dig,
                    "--samples-per-buffer", str(case["spb"]), "-T", str(case["thr"])]
            if case["filter"] is not None:
                args += ["-F", case["filter"]]
            csv = subprocess.run(args, check=True, capture_output=True, text=True).stdout
            pretty = subprocess.run([a if a != "csv" else "pretty" for a in args], check=True, capture_output=True,
                                    text=True).stdout
            first_bit, edges = O.parse_dig_csv(open(dig).read())
        drop_ts = dev["ts_mode"] != "none"
        rows = [r.split(",")[1:] if drop_ts else r.split(",") for r in csv.strip().splitlines()] if csv.strip() else []
        rec = dict(case)
        rec.update(n_samples=int(len(iq)), first_bit=int(first_bit), edges=[int(e) for e in edges], csv_rows=rows,
                   pretty=[l for l in pretty.splitlines() if "Decode Timestamp" not in l], sent=[m.hex() for m in msgs])
        out.append(rec)
        print("rx", case["name"], "edges", len(edges), "rows", max(len(rows) - 1, 0), "of", case["n_msgs"])
    json.dump(out, open(os.path.join(GOLD, "rx.json"), "w"), indent=0)


REC_CASES = [("nexa_awgn02_fs32", False), ("remote1_snr_dec4", False), ("remote1_spb1001", False), ("nexa_nofilter", False),
             ("nexa_spb5000", True), ("nexa_default_filter", False)]


def make_rec():
    """--rx-rec / --rx-rec-input recordings of the reference (src/ookiedokie.c:248-270): SHA-256, length and the first
    samples of the SC16Q11 file it writes."""
    import hashlib
    out = []
    by_name = {c["name"]: c for c in RX_CASES}
    for name, rec_input in REC_CASES:
        case = by_name[name]
        dev, iq, msgs = build_capture(case)
        with tempfile.TemporaryDirectory() as td:
            cap, rec = os.path.join(td, "c.sc16q11"), os.path.join(td, "rec.sc16q11")
            iq.tofile(cap)
            args = [REF, "--rx", "bladerf_file", "-A", cap, "-d", case["device"], "--rx-fmt", "csv", "-R", rec,
                    "--samples-per-buffer", str(case["spb"]), "-T", str(case["thr"])]
            if case["filter"] is not None:
                args += ["-F", case["filter"]]
            if rec_input:
                args += ["--rx-rec-input"]
            subprocess.run(args, check=True, capture_output=True, text=True)
            data = open(rec, "rb").read()
        x = np.frombuffer(data, dtype=np.int16).reshape(-1, 2)
        nz = int(np.flatnonzero(np.abs(x).sum(1) > 100)[0]) if np.any(np.abs(x).sum(1) > 100) else 0
        out.append(dict(name=name, rec_input=rec_input, n_samples=int(len(x)), sha256=hashlib.sha256(data).hexdigest(),
                        probe_at=nz, probe=x[nz:nz + 16].reshape(-1).tolist()))
        print("rec", name, rec_input, len(x))
    json.dump(out, open(os.path.join(GOLD, "rec.json"), "w"), indent=0)


if __name__ == "__main__":
    assert REF and FIR_TEST, "build oracle/_ref first (make -C oracle)"
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "rec":
        make_rec()
        sys.exit(0)
    make_fir()
    make_tx()
    make_rx()
    make_rec()
