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
SHM_SLOT_BYTES = 112 + 65536 * 56


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
        import platform
        # plain stores publish the payload before the sequence number: only on x86-64 (total store order)
        ok = os.path.isdir("/dev/shm") and platform.machine() in ("x86_64", "AMD64") and _same_host(world)
        flags = [None] * world
        dist.all_gather_object(flags, bool(ok))
        if all(flags):
            try:
                ex = _ShmExchange(rank, world, SHM_SLOT_BYTES)
                made = True
            except Exception:
                ex, made = None, False
            # the constructor's outcome is agreed on collectively too: one rank without the mapping would take the
            # collective path while the others spin on the slots
            flags = [None] * world
            dist.all_gather_object(flags, made)
            if not all(flags):
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
    if ex is not None and REC_BYTES + len(rec) * isz > SHM_SLOT_BYTES:
        # a rank-local fallback would leave the other ranks spinning on the shared-memory slots: refuse loudly instead
        raise RuntimeError(f"shared-memory stitch: {len(rec)} messages exceed the slot capacity "
                           f"({(SHM_SLOT_BYTES - REC_BYTES) // isz}); raise shard.SHM_SLOT_BYTES")
    if ex is not None:
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


# ---------------------------------------------------------------------------------------------------------------
# Stitch WITHOUT a per-step rendezvous (ranks of one host).
#
# stitch_and_gather() makes every rank wait, every step, until the slowest rank has published: the per-GPU decode
# takes the same ~1 ms everywhere, so what the lock-step adds is the jitter of N host processes.  Here a rank
# publishes its record for step s into slot s % RING of a shared-memory ring and goes on; the record of step s-1 is
# confirmed one step later, when everybody has normally long published it:
#
#   record  = state (1 provisional / 2 final) | exit carry | entry used | message count | messages
#   rank r's record for step s is FINAL as soon as the records of ranks 0..r for that step are consistent (each
#   entry used equals the predecessor's exit): no rank has to wait for anybody's confirmation, only for data.
#   Otherwise the lowest inconsistent rank re-runs its state-machine stage from its predecessor's (final) exit and
#   republishes; the ranks above it wait for their predecessor's FINAL record and do the same if needed.
#   Rank 0 assembles the message list of step s from the FINAL records of all ranks.
#
# A handle's tables must survive until its step is confirmed (resolve re-uses them), hence two handles used
# alternately and confirmation of step s-1 before step s+1 is enqueued (bench.py run_steps_multi).
# ---------------------------------------------------------------------------------------------------------------
class _ShmRing:
    HDR = 64
    RING = 4

    def __init__(self, rank, world, slot_bytes):
        import mmap
        import uuid
        self.rank, self.world, self.slot = rank, world, self.HDR + slot_bytes
        name = [f"/dev/shm/ookd_ring_{os.getpid()}_{uuid.uuid4().hex}" if rank == 0 else None]
        dist.broadcast_object_list(name, src=0)
        self.path = name[0]
        size = self.RING * world * self.slot
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(size)
        dist.barrier()
        self.f = open(self.path, "r+b")
        self.mm = mmap.mmap(self.f.fileno(), size)
        self.buf = np.frombuffer(self.mm, dtype=np.uint8)
        dist.barrier()
        if rank == 0:
            os.unlink(self.path)
        self.step = 0                        # steps published so far by this rank (same on every rank by construction)

    def slot_of(self, step, r):
        o = ((step % self.RING) * self.world + r) * self.slot
        return self.buf[o:o + self.slot]

    def publish(self, step, state, payload):
        sl = self.slot_of(step, self.rank)
        n = payload.size
        if n > self.slot - self.HDR:
            raise RuntimeError("shared-memory stitch: record exceeds the slot capacity; raise shard.SHM_SLOT_BYTES")
        sl[0:8].view(np.uint64)[0] = 0                               # invalidate while the payload changes
        sl[self.HDR:self.HDR + n] = payload
        sl[8:16].view(np.uint64)[0] = n
        sl[0:8].view(np.uint64)[0] = (step + 1) * 4 + state          # published (x86-64: stores are not reordered)

    def read(self, step, r, want_final, timeout_s=120.0):
        """Record of rank r for `step` (waits for it; with want_final for its FINAL version).  -> (state, bytes copy)."""
        import time
        sl = self.slot_of(step, r)
        seq = sl[0:8].view(np.uint64)
        t0, spins = None, 0
        while True:
            v = int(seq[0])
            if v // 4 == step + 1 and (v % 4 == 2 or (not want_final and v % 4 == 1)):
                ln = int(sl[8:16].view(np.uint64)[0])
                data = sl[self.HDR:self.HDR + ln].copy()
                if int(seq[0]) == v:                                  # not republished while we copied
                    return v % 4, data
                continue
            spins += 1
            if spins > 2000:
                if t0 is None:
                    t0 = time.perf_counter()
                elif time.perf_counter() - t0 > timeout_s:
                    raise RuntimeError(f"shared-memory stitch: rank {r} did not publish step {step}")
                time.sleep(0)


_ring = {}
_last_path = ["none"]


def stitch_description():
    """Which exchange the last multi-rank steps really used (recorded in the bench line)."""
    return _last_path[0]


def _shm_ring(rank, world):
    key = world
    if key not in _ring:
        import platform
        ok = os.path.isdir("/dev/shm") and platform.machine() in ("x86_64", "AMD64") and _same_host(world)
        flags = [None] * world
        dist.all_gather_object(flags, bool(ok))
        ring = None
        if all(flags):
            try:
                ring = _ShmRing(rank, world, SHM_SLOT_BYTES)
                made = True
            except Exception:
                ring, made = None, False
            flags = [None] * world
            dist.all_gather_object(flags, made)
            if not all(flags):
                ring = None
        _ring[key] = ring
    return _ring[key]


class PipelinedStitcher:
    """finish(runner, decoded) once per step, in step order; drain() before the results are used."""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world
        self.ring = _shm_ring(rank, world) if world > 1 else None
        self.pending = []                    # [step, runner, result, exit carry, entry used, final?]
        self.last_messages = None
        self.last = None
        self.resolves = 0
        _last_path[0] = ("shared-memory ring, no per-step rendezvous (records confirmed one step later); NCCL only for "
                         "the barriers and the timing all-reduce" if self.ring is not None else
                         "lock-step: one all-gather (NCCL) of carries + message blocks per step" if world > 1 else "n/a")

    @staticmethod
    def _payload(exit_c, entry_used, rec):
        hdr = carry_to_bytes(exit_c) + carry_to_bytes(entry_used) + np.array([0, len(rec), 0, 0], dtype=np.uint32).tobytes()
        return np.concatenate([np.frombuffer(hdr, dtype=np.uint8), rec.view(np.uint8).reshape(-1)])

    def confirm_pending(self):
        """Confirm every step published so far (normally the previous one, whose records everybody has long published).
        Must run before the handle of such a step is given new work: a wrong entry is repaired on its tables."""
        while self.pending:
            self._confirm(self.pending.pop(0))

    def publish(self, runner, decoded):
        """Publish this step's record without waiting for anybody.  -> (result, exit)."""
        res, exit_c = decoded if decoded is not None else runner.decode(None)
        if self.world == 1:
            self.last_messages = res["msgs_raw"]
            return res, exit_c
        if self.ring is None:
            res, exit_c, rounds, msgs = stitch_and_gather(runner, self.rank, self.world, decoded=(res, exit_c))
            self.last_messages = msgs
            return res, exit_c
        ring = self.ring
        step = ring.step
        ring.step += 1
        entry_used = tuple(res["entry_used"])
        ring.publish(step, 1, self._payload(exit_c, entry_used, res["msgs_raw"]))
        self.pending.append([step, runner, res, exit_c, entry_used])
        return res, exit_c

    def finish(self, runner, decoded, confirm_now=False):
        """confirm_pending() + publish() (+ immediate confirmation: needed when the step's handle is re-used at once).
        -> (result, exit, rounds, messages or None): the confirmed message list of a step appears in last_messages
        (rank 0) once that step has been confirmed."""
        self.confirm_pending()
        res, exit_c = self.publish(runner, decoded)
        if confirm_now:
            self.confirm_pending()
        return res, exit_c, 1, self.last_messages

    def _confirm(self, item):
        from .binding import MSG_DTYPE
        step, runner, res, exit_c, entry_used = item
        ring, rank, world = self.ring, self.rank, self.world
        if rank > 0:
            # records of ranks 0..rank-1 as published (provisional is enough while everything is consistent)
            recs = [ring.read(step, r, False)[1] for r in range(rank)]
            mine = carry_to_bytes(entry_used)
            ok = recs[rank - 1][:48].tobytes() == mine and all(
                recs[r][48:96].tobytes() == recs[r - 1][:48].tobytes() for r in range(1, rank))
            if not ok:
                # somebody below (or this rank) entered its shard in the wrong state: wait for the predecessor's FINAL
                # record and re-run the state-machine stage from its exit if that is not what was assumed
                _, pred = ring.read(step, rank - 1, True)
                true_entry = carry_from_bytes(pred[:48].tobytes())
                if true_entry != entry_used:
                    res, exit_c = runner.resolve(true_entry)
                    entry_used = true_entry
                    self.resolves += 1
        ring.publish(step, 2, self._payload(exit_c, entry_used, res["msgs_raw"]))
        self.last = (res, exit_c)
        if rank == 0:
            parts = [res["msgs_raw"].view(np.uint8).reshape(-1)]
            for r in range(1, world):
                _, d = ring.read(step, r, True)
                parts.append(d[REC_BYTES:])
            self.last_messages = np.concatenate(parts).view(MSG_DTYPE)

    def drain(self):
        while self.pending:
            self._confirm(self.pending.pop(0))
