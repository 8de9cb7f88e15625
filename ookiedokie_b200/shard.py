"""Time-sharded decode across ranks: one process per GPU, no collective on the data path.

Rank r owns input samples [r*S, (r+1)*S) of one continuous capture and reads `halo` samples of
history in front of its shard (FIR halo + one byte of decisions for the edge detector).  The only
exchange is the state-machine carry at each shard boundary (struct ookd_sm_carry, 48 bytes):

    1. every rank decodes its shard from a guessed entry state (rank 0: the true initial state);
    2. exits are all-gathered; a rank whose predecessor's exit differs from the entry it used
       re-runs ONLY its state-machine stage from the corrected entry (ookd_gpu_resolve; samples,
       decisions and edges are untouched);
    3. repeat until no rank changed -- at most world_size rounds, because rank 0's entry is exact
       and each pass fixes at least the first still-wrong boundary.

The fixed point equals the sequential decode (each shard is a deterministic function of its entry).
`runner` is anything with decode(entry) / resolve(entry) -> (result, exit_carry_tuple), so the
protocol itself is testable on CPU with gloo and a stand-in runner.
"""
import os

import numpy as np
import torch
import torch.distributed as dist

CARRY_BYTES = 16 + 32
MSG_REC = 8 + 8 + 4 + 32        # out_sample, buffer_idx, num_bits, data


def carry_to_bytes(c):
    state, k, num_bits, prev, data = c
    return np.array([state, k, num_bits, prev], dtype=np.uint32).tobytes() + bytes(data).ljust(32, b"\0")[:32]


def carry_from_bytes(b):
    b = bytes(b)
    s = np.frombuffer(b[:16], dtype=np.uint32)
    return (int(s[0]), int(s[1]), int(s[2]), int(s[3]), b[16:48])


INITIAL_CARRY = (0, 0, 0, 0, bytes(32))


def _dev():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


REC_BYTES = 112         # exit carry (48) | entry used (48) | changed flag (4) | message count (4) | pad (8)


class _StitchBuffers:
    """Pinned host and device staging reused across steps (one collective, one synchronisation per step)."""

    def __init__(self, world, cap, isz, dev):
        self.cap = cap
        self.nbytes = REC_BYTES + cap * isz
        pin = dev.type == "cuda"
        self.h_mine = torch.zeros(self.nbytes, dtype=torch.uint8, pin_memory=pin)
        self.h_all = torch.zeros(world * self.nbytes, dtype=torch.uint8, pin_memory=pin)
        self.d_mine = torch.zeros(self.nbytes, dtype=torch.uint8, device=dev)
        self.d_all = torch.zeros(world * self.nbytes, dtype=torch.uint8, device=dev)
        self.np_mine = self.h_mine.numpy()
        self.np_all = self.h_all.numpy()


_stitch_cache = {}
_stitch_cap = {}


class _ShmExchange:
    """All-gather of small per-rank byte blocks through a shared-memory file (ranks of ONE host): the carry
    records and message lists are a few hundred KB at most and every rank already has them in host memory,
    so a host-mediated exchange beats staging them through the GPUs and a collective (~20 us against ~300 us
    per step).  Slot layout per (parity, rank): u64 sequence number | u64 length | payload.  A rank publishes
    step s by writing the payload and then the sequence number; readers spin on the sequence numbers.  Two
    parities: nobody can be two steps ahead of a rank that is still reading, because publishing step s+1
    happens after reading step s."""

    HDR = 64

    def __init__(self, rank, world, slot_bytes):
        import mmap
        import uuid
        self.rank, self.world, self.slot = rank, world, self.HDR + slot_bytes
        name = [f"/dev/shm/ookd_stitch_{os.getpid()}_{uuid.uuid4().hex}" if rank == 0 else None]
        dist.broadcast_object_list(name, src=0)
        self.path = name[0]
        size = 2 * world * self.slot
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(size)
        dist.barrier()
        self.f = open(self.path, "r+b")
        self.mm = mmap.mmap(self.f.fileno(), size)
        self.buf = np.frombuffer(self.mm, dtype=np.uint8)
        self.seq = 0
        dist.barrier()
        if rank == 0:
            os.unlink(self.path)          # the mapping stays valid; nothing is left behind on exit

    def _slot(self, parity, r):
        o = (parity * self.world + r) * self.slot
        return self.buf[o:o + self.slot]

    def exchange(self, payload, timeout_s=120.0):
        """payload: uint8 array (<= slot capacity).  -> list of uint8 arrays, one per rank (copies)."""
        import time
        self.seq += 1
        par = self.seq & 1
        mine = self._slot(par, self.rank)
        n = payload.size
        mine[self.HDR:self.HDR + n] = payload
        mine[8:16].view(np.uint64)[0] = n
        mine[0:8].view(np.uint64)[0] = self.seq          # published (x86 stores are not reordered)
        out = []
        t0 = None
        for r in range(self.world):
            sl = self._slot(par, r)
            seq = sl[0:8].view(np.uint64)
            spins = 0
            while int(seq[0]) != self.seq:
                spins += 1
                if spins > 2000:
                    if t0 is None:
                        t0 = time.perf_counter()
                    elif time.perf_counter() - t0 > timeout_s:
                        raise RuntimeError(f"shared-memory stitch: rank {r} did not publish step {self.seq}")
                    time.sleep(0)
            ln = int(sl[8:16].view(np.uint64)[0])
            out.append(sl[self.HDR:self.HDR + ln].copy())
        return out


_shm = {}
SHM_SLOT_BYTES = 112 + 16384 * 56


def _same_host(world):
    names = [None] * world
    import socket
    dist.all_gather_object(names, socket.gethostname())
    return all(n == names[0] for n in names)


def _shm_exchange(rank, world):
    """The shared-memory exchange of this process group, or None when the ranks span hosts / it cannot be set up."""
    key = world
    if key not in _shm:
        ex = None
        try:
            ok = os.path.isdir("/dev/shm") and _same_host(world)
            flags = [None] * world
            dist.all_gather_object(flags, bool(ok))
            if all(flags):
                ex = _ShmExchange(rank, world, SHM_SLOT_BYTES)
        except Exception:
            ex = None
        _shm[key] = ex
    return _shm[key]


def stitch_and_gather(runner, rank, world, msg_cap=None, decoded=None):
    """stitch() + gather_messages_raw() with ONE collective in the common case: every rank ships its
    carry record together with its (padded) message block; if the records are consistent the job is done.
    The block is sized from the largest per-rank message count of the previous call (every rank saw the same
    counts, so every rank picks the same size).  Returns (result, exit, rounds, messages on rank 0 | None)."""
    from .binding import MSG_DTYPE
    # decoded = (result, exit) of runner.decode(None) when the caller has already waited for the decode (so that it
    # could enqueue the next window before the cross-rank exchange of this one)
    res, exit_c = decoded if decoded is not None else runner.decode(None)
    if world == 1:
        return res, exit_c, 1, res["msgs_raw"]
    entry_used = tuple(res["entry_used"])
    dev = _dev()
    isz = MSG_DTYPE.itemsize
    rec = res["msgs_raw"]
    ex = _shm_exchange(rank, world) if msg_cap is None else None
    if ex is not None and REC_BYTES + len(rec) * isz <= SHM_SLOT_BYTES:
        # ranks of one host: records and message lists go through shared memory, no device round trip
        hdr = carry_to_bytes(exit_c) + carry_to_bytes(entry_used) + np.array([0, len(rec), 0, 0], dtype=np.uint32).tobytes()
        payload = np.concatenate([np.frombuffer(hdr, dtype=np.uint8), rec.view(np.uint8).reshape(-1)])
        parts = ex.exchange(payload)
        ok = all(parts[r][48:96].tobytes() == parts[r - 1][:48].tobytes() for r in range(1, world))
        if ok:
            msgs = None
            if rank == 0:
                msgs = np.concatenate([p[REC_BYTES:] for p in parts]).view(MSG_DTYPE)     # (bytes first: concatenating
                                                                                          # structured arrays is slow)
            return res, exit_c, 1, msgs
        res, exit_c, rounds = _stitch_rounds(runner, rank, world, res, exit_c, entry_used, dev)
        msgs = gather_messages_raw(res["msgs_raw"], rank, world, res.get("_counts"))
        return res, exit_c, rounds + 1, msgs
    cap = msg_cap if msg_cap is not None else _stitch_cap.get(world, 1024)
    key = (world, cap, dev.type, dev.index)
    sb = _stitch_cache.get(key)
    if sb is None:
        sb = _stitch_cache[key] = _StitchBuffers(world, cap, isz, dev)
    n_ship = min(len(rec), cap)
    hdr = carry_to_bytes(exit_c) + carry_to_bytes(entry_used) + np.array([0, len(rec), 0, 0], dtype=np.uint32).tobytes()
    sb.np_mine[:REC_BYTES] = np.frombuffer(hdr, dtype=np.uint8)
    if n_ship:
        sb.np_mine[REC_BYTES:REC_BYTES + n_ship * isz] = rec[:n_ship].view(np.uint8).reshape(-1)
    sb.d_mine.copy_(sb.h_mine, non_blocking=True)
    dist.all_gather_into_tensor(sb.d_all, sb.d_mine)
    sb.h_all.copy_(sb.d_all, non_blocking=True)
    if dev.type == "cuda":
        torch.cuda.current_stream().synchronize()
    g = sb.np_all.reshape(world, sb.nbytes)
    counts = g[:, 100:104].copy().view(np.uint32).reshape(-1).tolist()
    if msg_cap is None:
        want = 1024
        while want < max(counts) + max(counts) // 4 + 64:
            want *= 2
        _stitch_cap[world] = want                      # same on every rank: derived from the gathered counts
    ok = all(g[r, 48:96].tobytes() == g[r - 1, :48].tobytes() for r in range(1, world))
    if ok and max(counts) <= cap:
        msgs = None
        if rank == 0:
            msgs = np.concatenate([g[r, REC_BYTES:REC_BYTES + counts[r] * isz] for r in range(world)]).view(MSG_DTYPE)
        return res, exit_c, 1, msgs
    if not ok:
        # inconsistent entries: the round protocol re-runs the state-machine stage of the shards that guessed wrong
        res, exit_c, rounds = _stitch_rounds(runner, rank, world, res, exit_c, entry_used, dev)
        msgs = gather_messages_raw(res["msgs_raw"], rank, world, res.get("_counts"))
        return res, exit_c, rounds + 1, msgs
    # consistent, but a message list did not fit the block (first call, or a burst): fetch the lists separately
    msgs = gather_messages_raw(res["msgs_raw"], rank, world, counts)
    return res, exit_c, 1, msgs


def stitch(runner, rank, world, guess=None):
    """Run the protocol above.  Returns (result, exit, rounds).  One 112-byte all-gather per round; the
    last round's records also carry every rank's message count (used by gather_messages_raw).

    With handles created with sm_warmup the first decode already enters each shard in the state its own
    warm-up history leads to (result["entry_used"]); the first gather then only CONFIRMS that every
    entry equals the predecessor's exit, and no rank runs a second pass."""
    entry_used = None if rank == 0 else guess
    res, exit_c = runner.decode(entry_used)
    if world == 1:
        return res, exit_c, 1
    entry_used = tuple(res["entry_used"]) if "entry_used" in res else (entry_used or INITIAL_CARRY)
    return _stitch_rounds(runner, rank, world, res, exit_c, entry_used, _dev())


def _stitch_rounds(runner, rank, world, res, exit_c, entry_used, dev):
    rounds = 1
    changed_last = 0
    while True:
        n_msgs = len(res["msgs_raw"]) if "msgs_raw" in res else len(res["msgs"])
        rec = (carry_to_bytes(exit_c) + carry_to_bytes(entry_used) +
               np.array([changed_last, n_msgs, 0, 0], dtype=np.uint32).tobytes())
        mine = torch.frombuffer(bytearray(rec), dtype=torch.uint8).to(dev)
        gathered = torch.empty(world * REC_BYTES, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(gathered, mine)
        g = gathered.cpu().numpy().reshape(world, REC_BYTES)
        # consistent iff every rank's entry is its predecessor's exit (every rank sees the same records)
        ok = all(g[r, 48:96].tobytes() == g[r - 1, :48].tobytes() for r in range(1, world))
        if ok:
            res["_counts"] = g[:, 100:104].copy().view(np.uint32).reshape(-1).tolist()
            return res, exit_c, rounds
        changed_last = 0
        if rank > 0:
            true_entry = carry_from_bytes(g[rank - 1, :48].tobytes())
            if true_entry != entry_used:
                res, exit_c = runner.resolve(true_entry)
                entry_used = true_entry
                changed_last = 1
        rounds += 1
        if rounds > world + 3:
            raise RuntimeError("shard stitch did not converge")


def gather_messages(msgs, rank, world, nbytes):
    """msgs: [(out_sample, buffer_idx, num_bits, data)] per rank -> full ordered list on rank 0 (None elsewhere)."""
    from .binding import MSG_DTYPE, msgs_to_tuples
    rec = np.zeros(len(msgs), dtype=MSG_DTYPE)
    for i, (o, b, nb, data) in enumerate(msgs):
        rec[i]["out_sample"], rec[i]["buffer_idx"], rec[i]["num_bits"] = o, b, nb
        rec[i]["data"][:len(data)] = np.frombuffer(bytes(data), dtype=np.uint8)
    out = gather_messages_raw(rec, rank, world)
    return None if out is None else msgs_to_tuples(out, nbytes)


def gather_messages_raw(rec, rank, world, counts=None):
    """Structured message arrays (binding.MSG_DTYPE) per rank -> concatenated array on rank 0 (None elsewhere).
    Two small collectives: the counts, then one padded byte block per rank."""
    if world == 1:
        return rec
    dev = _dev()
    if counts is None:
        cnt = torch.tensor([len(rec)], dtype=torch.int64, device=dev)
        cl = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(cl, cnt)
        counts = [int(c) for c in torch.cat(cl).cpu().tolist()]
    cap = max(max(counts), 1)
    isz = rec.dtype.itemsize
    buf = np.zeros(cap * isz, dtype=np.uint8)
    buf[:len(rec) * isz] = rec.view(np.uint8).reshape(-1)
    mine = torch.from_numpy(buf).to(dev)
    allb = torch.empty(world * cap * isz, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allb, mine)
    if rank != 0:
        return None
    flat = allb.cpu().numpy()
    parts = [flat[r * cap * isz: r * cap * isz + counts[r] * isz].view(rec.dtype) for r in range(world)]
    return np.concatenate(parts)


class GpuShardRunner:
    """runner over one ookiedokie_b200.binding.Gpu handle and a resident (device or pinned host) shard."""

    def __init__(self, gpu, iq, first_sample, n_samples, last):
        self.gpu, self.iq, self.first, self.n, self.last = gpu, iq, first_sample, n_samples, last
        self.launches = 0
        self.fir_ms = 0.0
        self.screen_ms = 0.0
        self.kernel_ms = 0.0
        self.host_syncs = 0

    def _acc(self, res):
        self.launches += res["gpu_launches"]
        self.fir_ms += res["fir_ms"]
        self.screen_ms += res.get("screen_ms", 0.0)
        self.host_syncs += res.get("host_syncs", 0)
        self.kernel_ms += res["kernel_ms"]

    def begin(self, entry=None):
        """Enqueue the decode now; the next decode(entry) call only waits for it."""
        self.gpu.decode_begin(self.iq, self.first, self.n, self.last, entry)
        self._begun = True

    def decode(self, entry):
        if getattr(self, "_begun", False):
            self._begun = False
            res, ex = self.gpu.decode_end()
        else:
            res, ex = self.gpu.decode_shard(self.iq, self.first, self.n, self.last, entry)
        self._acc(res)
        return res, ex

    def resolve(self, entry):
        res, ex = self.gpu.resolve(entry)
        self._acc(res)
        return res, ex
