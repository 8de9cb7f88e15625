"""Worker for tests/test_shard_gloo.py: one rank of the time-shard stitch protocol on CPU (gloo).
The GPU runner is replaced by a stand-in that steps the CPU oracle's state machine over the shard's
threshold decisions, buffer by buffer with device_process's drop rule -- the protocol code
(ookiedokie_b200/shard.py) is the real one."""
import json
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
from ookiedokie_b200 import shard as S  # noqa: E402
import ookd_testutil as util  # noqa: E402


class OracleShardRunner:
    def __init__(self, device, sample_rate, bits, out_lo, opb):
        self.device, self.rate, self.bits, self.out_lo, self.opb = device, sample_rate, bits, out_lo, opb
        self.calls = []

    def _run(self, entry):
        sm = O.Sm(self.device, self.rate)
        if entry is not None:
            sm.set_state(entry)
        msgs = []
        nbytes = (self.device["num_bits"] + 7) // 8
        for b0 in range(0, len(self.bits), self.opb):
            buf = self.bits[b0:b0 + self.opb]
            total, r = 0, 0
            while total < len(buf) and r != -1:
                r, n = sm.process(buf[total:])
                total += n
                if r == 1:
                    m = self.out_lo + b0 + total - 1
                    msgs.append((m, (m + 1 - 1) // self.opb, sm.get_state()[2], sm.data()[:nbytes]))
        from ookiedokie_b200.binding import MSG_DTYPE
        raw = np.zeros(len(msgs), dtype=MSG_DTYPE)
        for i, (m, b, nb, data) in enumerate(msgs):
            raw[i]["out_sample"], raw[i]["buffer_idx"], raw[i]["num_bits"] = m, b, nb
            raw[i]["data"][:len(data)] = np.frombuffer(bytes(data), dtype=np.uint8)
        return dict(msgs=msgs, msgs_raw=raw, entry_used=entry if entry is not None else S.INITIAL_CARRY), sm.get_state()

    def decode(self, entry):
        self.calls.append("decode")
        return self._run(entry)

    def resolve(self, entry):
        self.calls.append("resolve")
        return self._run(entry)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    out_path = sys.argv[1]
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spb = 8192
    dev = O.load_device("p3l-nexa2012")
    iq, sent, _ = util.capture(dev, 5, sigma=0.02, phase=0.4, seed=77, fields=util.nexa_fields)
    stages = O.load_filter("fs32_fs4")
    full = O.rx(iq, stages, dev, samples_per_buffer=spb, want_bits=True)
    n_buf = full["n_buffers"]
    per = (n_buf + world - 1) // world
    lo, hi = rank * per * spb, min((rank + 1) * per * spb, len(full["bits"]))
    runner = OracleShardRunner(dev, 3000000, full["bits"][lo:hi], lo, spb)
    nbytes = (dev["num_bits"] + 7) // 8
    if len(sys.argv) > 2 and sys.argv[2] == "fused":
        # the single-collective form bench.py uses; twice, so that the adaptive block size is exercised, and once
        # with a block too small for the message lists (falls back to the separate gather)
        from ookiedokie_b200.binding import msgs_to_tuples
        for cap in (None, None, 1, 4096):
            res, exit_c, rounds, raw = S.stitch_and_gather(runner, rank, world, msg_cap=cap)
            msgs = msgs_to_tuples(raw, nbytes) if rank == 0 else None
            if rank == 0:
                assert [tuple(m) for m in msgs] == [tuple(m) for m in full["msgs"]], cap
    elif len(sys.argv) > 2 and sys.argv[2] == "pipelined":
        # the rendezvous-free form bench.py uses at N > 1: records published into the shared-memory ring, each step
        # confirmed one step later; more steps than ring slots, fresh runner per step (as handles alternate)
        from ookiedokie_b200.binding import msgs_to_tuples
        st = S.PipelinedStitcher(rank, world)
        assert st.ring is not None, "shared-memory ring unavailable"
        seen = 0
        for step in range(7):
            r = OracleShardRunner(dev, 3000000, full["bits"][lo:hi], lo, spb)
            decoded = r.decode(None)
            st.confirm_pending()
            if rank == 0 and step > 0:
                assert [tuple(m) for m in msgs_to_tuples(st.last_messages, nbytes)] == [tuple(m) for m in full["msgs"]], step
                seen += 1
            st.publish(r, decoded)
            runner.calls += r.calls
        st.drain()
        rounds = 1
        msgs = msgs_to_tuples(st.last_messages, nbytes) if rank == 0 else None
        if rank == 0:
            assert seen == 6
        # a second stitcher on the same ring continues the step numbering
        st2 = S.PipelinedStitcher(rank, world)
        r = OracleShardRunner(dev, 3000000, full["bits"][lo:hi], lo, spb)
        st2.finish(r, r.decode(None), confirm_now=True)
        if rank == 0:
            assert [tuple(m) for m in msgs_to_tuples(st2.last_messages, nbytes)] == [tuple(m) for m in full["msgs"]]
    else:
        res, exit_c, rounds = S.stitch(runner, rank, world)
        msgs = S.gather_messages(res["msgs"], rank, world, nbytes)
    if rank == 0:
        ok = [tuple(m) for m in msgs] == [tuple(m) for m in full["msgs"]]
        json.dump(dict(ok=ok, n=len(msgs), want=len(full["msgs"]), rounds=rounds), open(out_path, "w"))
    json.dump(dict(calls=runner.calls), open(out_path + f".rank{rank}", "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
