"""Time shards on one GPU: FIR halo, edge continuity and state-machine carry across shard boundaries."""
import numpy as np
import pytest

from oracle import oracle as O
from ookiedokie_b200 import binding as B
from ookiedokie_b200 import shard as S
import ookd_testutil as util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("devname,filt,spb,n_shards", [
    ("p3l-nexa2012", "fs32_fs4", 8192, 3),
    ("p3l-nexa2012", "fs32_fs4", 8192, 7),
    ("p3l-nexa2012", "fs128_fs16_dec4", 8192, 4),
    ("unknown-remote1", "fs128_fs16_dec4", 1001, 3),
    ("p3l-nexa2012", "fs64_fs8", 4096, 5),
])
def test_shards_equal_whole_capture(devname, filt, spb, n_shards):
    dev = O.load_device(devname)
    fields = util.nexa_fields if "nexa" in devname else util.remote_fields
    iq, msgs, _ = util.capture(dev, 4, sigma=0.02, amplitude=0.9, phase=1.1, seed=5, fields=fields)
    stages = O.load_filter(filt)
    sm = util.sm_spec(dev, stages)
    ref = O.rx(iq, stages, dev, samples_per_buffer=spb)
    g = B.Gpu(filter_stages=stages, sm=sm, samples_per_buffer=spb, sm_chunk_buffers=5)
    whole = g.decode(iq)
    assert whole["msgs"] == ref["msgs"]
    dec, halo = g.total_decimation, g.halo
    align = np.lcm(spb, dec)
    n = len(iq)
    per = max(align, (n // n_shards) // align * align)
    bounds = list(range(0, n, per))[:n_shards] + [n]
    # (a) sequential: each shard entered with its predecessor's exit
    carry, all_msgs, all_edges = None, [], []
    exits = []
    for i in range(len(bounds) - 1):
        first, cnt = bounds[i], bounds[i + 1] - bounds[i]
        h = min(halo, first)
        res, carry = g.decode_shard(iq[first - h: first + cnt], first, cnt, i == len(bounds) - 2, carry)
        all_msgs += res["msgs"]
        all_edges.append(g.edges()[1])
        exits.append(carry)
    assert all_msgs == ref["msgs"]
    assert np.array_equal(np.concatenate(all_edges), ref["edges"])
    # (b) every shard from a guessed entry, then resolve with the true one: same exits and messages
    for i in range(1, len(bounds) - 1):
        first, cnt = bounds[i], bounds[i + 1] - bounds[i]
        h = min(halo, first)
        res_g, exit_g = g.decode_shard(iq[first - h: first + cnt], first, cnt, i == len(bounds) - 2, S.INITIAL_CARRY)
        res_t, exit_t = g.resolve(exits[i - 1])
        assert exit_t == exits[i]
        want = [m for m in ref["msgs"] if bounds[i] // dec <= m[0] < bounds[i + 1] // dec or
                (i == len(bounds) - 2 and m[0] >= bounds[i] // dec)]
        assert res_t["msgs"] == want


@pytest.mark.parametrize("devname,filt,spb", [("p3l-nexa2012", "fs32_fs4", 8192), ("unknown-remote1", "fs128_fs16_dec4", 8192)])
def test_warmup_history_gives_the_true_entry(devname, filt, spb):
    """sm_warmup: a shard decoded WITHOUT an entry reads one chunk of history and enters in the state the
    sequential run has there, so its messages/edges are already final and entry_used == predecessor's exit."""
    dev = O.load_device(devname)
    fields = util.nexa_fields if "nexa" in devname else util.remote_fields
    iq, msgs, _ = util.capture(dev, 10, sigma=0.02, amplitude=0.9, phase=0.3, seed=9, fields=fields, lead=12000)
    stages = O.load_filter(filt)
    sm = util.sm_spec(dev, stages)
    ref = O.rx(iq, stages, dev, samples_per_buffer=spb)
    g = B.Gpu(filter_stages=stages, sm=sm, samples_per_buffer=spb, sm_chunk_buffers=56, sm_warmup=1)
    dec, halo = g.total_decimation, g.halo
    assert halo >= 56 * spb          # one chunk of history, longer than a message
    align = np.lcm(spb, dec)
    n = len(iq)
    per = (n // 3) // align * align
    bounds = [0, per, 2 * per, n]
    assert per >= halo
    prev_exit, all_msgs, all_edges = None, [], []
    for i in range(3):
        first, cnt = bounds[i], bounds[i + 1] - bounds[i]
        h = min(halo, first)
        res, ex = g.decode_shard(iq[first - h: first + cnt], first, cnt, i == 2, None)
        if i == 0:
            assert res["entry_is_provisional"] == 0
        else:
            assert res["entry_is_provisional"] == 1
            assert res["entry_used"] == prev_exit
        fb, e = g.edges()
        all_edges.append(e)
        assert len(g.bits()) == res["n_out"]
        all_msgs += res["msgs"]
        prev_exit = ex
    assert all_msgs == ref["msgs"]
    assert np.array_equal(np.concatenate(all_edges), ref["edges"])
    # an explicit (corrected) entry after a warm decode bypasses the warm-up chunk
    res_t, ex_t = g.resolve(prev_exit if False else res["entry_used"])
    assert res_t["msgs"] == res["msgs"] and ex_t == ex
