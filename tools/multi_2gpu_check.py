"""On a box with >= 2 GPUs: ookd_gpu_multi_* with DISTINCT devices (one host thread per device) against the oracle, host
and device-resident input, and the CLI's --gpus 2 against --gpus 1."""
import os, subprocess, sys, tempfile
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import numpy as np, torch
from oracle import oracle as O
from ookiedokie_b200 import binding as B
import ookd_testutil as util

assert torch.cuda.device_count() >= 2, "needs two GPUs"
for devname, filt, spb in (("p3l-nexa2012", "fs32_fs4", 8192), ("unknown-remote1", "fs128_fs16_dec4", 1001)):
    dev = O.load_device(devname)
    fields = util.nexa_fields if "nexa" in devname else util.remote_fields
    iq, msgs, _ = util.capture(dev, 12, sigma=0.03, amplitude=0.8, phase=0.4, seed=21, fields=fields, glitches=((9000, 100),))
    stages = O.load_filter(filt)
    sm = util.sm_spec(dev, stages)
    ref = O.rx(iq, stages, dev, samples_per_buffer=spb)
    for ids in ([0, 1], [1, 0], [0, 1, 0, 1]):
        m = B.MultiGpu(ids, filter_stages=stages, sm=sm, samples_per_buffer=spb, sm_chunk_buffers=5)
        got, _ = m.decode(iq)
        assert got["msgs"] == ref["msgs"], (devname, ids)
        fb, edges = m.edges()
        assert fb == ref["first_bit"] and np.array_equal(edges, ref["edges"]), (devname, ids)
        # device-resident shards, each on its own GPU
        halo, n = m.halo, len(iq)
        flat = np.ascontiguousarray(iq).reshape(-1)
        bufs, ptrs = [], []
        for g, d in enumerate(ids):
            first, cnt = m.shard_range(0, n, g)
            h = min(halo, first)
            t = torch.from_numpy(flat[2 * (first - h): 2 * (first + cnt)].copy()).to(f"cuda:{d}")
            bufs.append(t)
            ptrs.append(t.data_ptr())
        got_d, _ = m.decode(None, 0, n, last=True, device_ptrs=ptrs)
        assert got_d["msgs"] == ref["msgs"], (devname, ids, "device")
        m.close()
        print("ok", devname, ids, len(ref["msgs"]), "messages", flush=True)

# CLI
dev = O.load_device("p3l-nexa2012")
iq, msgs, _ = util.capture(dev, 20, sigma=0.02, amplitude=0.9, phase=0.2, seed=3, fields=util.nexa_fields)
with tempfile.TemporaryDirectory() as td:
    path = os.path.join(td, "c.sc16q11")
    np.ascontiguousarray(iq).astype("<i2").tofile(path)
    exe = os.path.join("ookiedokie_b200", "bin", "ookiedokie-b200")
    env = dict(os.environ, OOKD_DATA_DIR=os.path.abspath("ookiedokie_b200/data") + "/", OOKD_RX_WINDOW=str(1 << 20))
    outs = []
    for extra in ([], ["--gpus", "2"], ["--gpu-ids", "1,0"]):
        r = subprocess.run([exe, "--rx", "bladerf_file", "-A", path, "-d", "p3l-nexa2012", "-F", "fs32_fs4", "--rx-fmt", "csv", "--window", "1048576"] + extra,
                           capture_output=True, text=True, env=env, timeout=120)
        assert r.returncode == 0, r.stderr
        outs.append([l.split(",")[1:] for l in r.stdout.strip().splitlines()])      # (column 0 is the decode timestamp)
    assert outs[0] == outs[1] == outs[2] and len(outs[0]) >= 10, [len(o) for o in outs]
    print("ok CLI --gpus 2 / --gpu-ids 1,0:", len(outs[0]), "rows identical")
