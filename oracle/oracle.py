"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/_ref/libookd_oracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product never does.

Besides the bindings it restates, in Python, the two *loaders* the reference
runs before the hot loop so that the oracle can be driven from the same JSON
files as the product:

* filter JSON  -> stage list            (reference src/fir.c:68-249)
* device JSON  -> state/trigger tables  (reference src/device.c:76-254 and the
  slot-assignment rule of get_or_reserve_state, src/state_machine.c:206-247)

and the transmit-side sample generator used to build synthetic captures
(reference src/state_machine.c:565-873, src/formatter.c:140-255,755-846).
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
DATA = os.path.join(REPO, "ookiedokie_b200", "data")
REF_DIR = os.path.join(HERE, "_ref")
MSG_BYTES = 32

COND = {"always": 1, "pulse_start": 2, "pulse_end": 3, "timeout": 4, "msg_complete": 5}
ACT = {"none": 1, "append_0": 2, "append_1": 3, "output_data": 4}


def build(ref=True):
    """(Re)build the oracle library (and the reference, when its sources exist)."""
    target = "all" if ref else "oracle"
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


class _Msg(C.Structure):
    _fields_ = [("out_sample", C.c_uint64), ("buffer_idx", C.c_uint64),
                ("num_bits", C.c_uint32), ("data", C.c_uint8 * MSG_BYTES)]


class _RxResult(C.Structure):
    _fields_ = [("n_out", C.c_uint64), ("n_buffers", C.c_uint64),
                ("filtered", C.POINTER(C.c_float)), ("bits", C.POINTER(C.c_uint8)),
                ("first_bit", C.c_uint8), ("n_edges", C.c_uint64),
                ("edges", C.POINTER(C.c_uint64)), ("n_msgs", C.c_uint64),
                ("msgs", C.POINTER(_Msg))]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(REF_DIR, "libookd_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.ookd_oracle_fir_create.restype = C.c_void_p
        L.ookd_oracle_fir_create.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ookd_oracle_fir_run.restype = C.c_size_t
        L.ookd_oracle_fir_run.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.ookd_oracle_fir_reset.argtypes = [C.c_void_p]
        L.ookd_oracle_fir_destroy.argtypes = [C.c_void_p]
        L.ookd_oracle_fir_total_decimation.restype = C.c_uint32
        L.ookd_oracle_fir_total_decimation.argtypes = [C.c_void_p]
        L.ookd_oracle_threshold.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_void_p]
        L.ookd_oracle_sc16q11_to_cf.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.ookd_oracle_sm_create.restype = C.c_void_p
        L.ookd_oracle_sm_create.argtypes = [C.c_uint32] + [C.c_void_p] * 7 + [C.c_uint32, C.c_uint32]
        L.ookd_oracle_sm_destroy.argtypes = [C.c_void_p]
        L.ookd_oracle_sm_process.restype = C.c_int
        L.ookd_oracle_sm_process.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        L.ookd_oracle_sm_data.restype = C.POINTER(C.c_uint8)
        L.ookd_oracle_sm_data.argtypes = [C.c_void_p]
        L.ookd_oracle_sm_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.ookd_oracle_sm_set_state.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        L.ookd_oracle_sm_num_bits.restype = C.c_uint32
        L.ookd_oracle_sm_num_bits.argtypes = [C.c_void_p]
        L.ookd_oracle_rx.restype = C.c_int
        L.ookd_oracle_rx.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_float,
                                     C.c_uint32, C.c_int, C.c_int, C.POINTER(_RxResult)]
        L.ookd_oracle_rx_free.argtypes = [C.POINTER(_RxResult)]
        L.ookd_oracle_synth.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64,
                                        C.c_int32, C.c_int32, C.c_int32, C.c_uint64]
        L.ookd_oracle_synth_ex.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64,
                                           C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.c_uint32]
        _lib = L
    return _lib


# ----------------------------------------------------------------------------
# Loaders
# ----------------------------------------------------------------------------
def find_data(kind, name):
    """kind in {'filters','devices'}; accepts a path or a bare name."""
    for cand in (name, name + ".json", os.path.join(DATA, kind, name + ".json")):
        if os.path.isfile(cand):
            return cand
    raise FileNotFoundError(f"{kind}/{name}")


def load_filter(name):
    """-> list of (decimation, float32 taps).  reference src/fir.c:118-226."""
    root = json.load(open(find_data("filters", name)))
    stages = []
    for st in root["filter"]["stages"]:
        dec = int(st.get("decimation", 1))
        taps = np.array([np.float32(float(t)) for t in st["taps"]], dtype=np.float32)
        stages.append((dec, taps))
    return stages


def filter_total_decimation(stages):
    d = 1
    for dec, _ in stages:
        d *= dec
    return d


def load_device(name):
    """-> dict(name, num_bits, states[...], fields[...], ts_mode) with the
    reference's state-slot assignment (state_machine.c:206-247)."""
    dev = json.load(open(find_data("devices", name)))["device"]
    jstates = dev["states"]
    n = len(jstates)
    slots = [None] * n

    def slot_of(nm):
        if nm.lower() == "reset" and slots[0] is None:
            slots[0] = nm
            return 0
        for i in range(n):
            if slots[i] is None:
                slots[i] = nm
                return i
            if slots[i] == nm:
                return i
        raise ValueError(f"no room for state {nm}")

    states = [None] * n
    for js in jstates:
        idx = slot_of(js["name"])
        trigs = []
        for jt in js["triggers"]:
            trigs.append(dict(cond=COND[jt["condition"].lower()],
                              duration_us=int(jt["duration_us"]) if isinstance(jt.get("duration_us"), int) else 0,
                              action=ACT[jt["action"].lower()] if isinstance(jt.get("action"), str) else ACT["none"],
                              next=slot_of(jt["state"])))
        states[idx] = dict(name=js["name"],
                           duration_us=int(js["duration_us"]) if isinstance(js.get("duration_us"), int) else 0,
                           timeout_us=int(js["timeout_us"]) if isinstance(js.get("timeout_us"), int) else 0,
                           triggers=trigs)
    return dict(name=dev["name"], num_bits=int(dev["num_bits"]), states=states,
                fields=dev["fields"], ts_mode=dev.get("ts_mode", "none"))


# ----------------------------------------------------------------------------
# Object wrappers
# ----------------------------------------------------------------------------
class Fir:
    def __init__(self, stages):
        self.stages = stages
        dec = np.array([s[0] for s in stages], dtype=np.uint32)
        nt = np.array([len(s[1]) for s in stages], dtype=np.uint32)
        taps = np.concatenate([s[1] for s in stages]).astype(np.float32)
        self.h = lib().ookd_oracle_fir_create(len(stages), dec.ctypes.data, nt.ctypes.data, taps.ctypes.data)
        if not self.h:
            raise ValueError("bad filter")
        self.total_decimation = int(lib().ookd_oracle_fir_total_decimation(self.h))

    def reset(self):
        lib().ookd_oracle_fir_reset(self.h)

    def run(self, iq_f32):
        """iq_f32: float32 array of shape (n, 2) -> (m, 2)."""
        x = np.ascontiguousarray(iq_f32, dtype=np.float32).reshape(-1, 2)
        out = np.empty((x.shape[0] + 1, 2), dtype=np.float32)
        m = lib().ookd_oracle_fir_run(self.h, x.ctypes.data, x.shape[0], out.ctypes.data)
        return out[:m].copy()

    def __del__(self):
        if getattr(self, "h", None):
            lib().ookd_oracle_fir_destroy(self.h)
            self.h = None


class Sm:
    def __init__(self, device, sample_rate):
        st = device["states"]
        dur = np.array([s["duration_us"] for s in st], dtype=np.uint64)
        tmo = np.array([s["timeout_us"] for s in st], dtype=np.uint64)
        off = np.zeros(len(st) + 1, dtype=np.uint32)
        cond, tdur, act, nxt = [], [], [], []
        for i, s in enumerate(st):
            for t in s["triggers"]:
                cond.append(t["cond"]); tdur.append(t["duration_us"]); act.append(t["action"]); nxt.append(t["next"])
            off[i + 1] = len(cond)
        cond = np.array(cond, dtype=np.int32); tdur = np.array(tdur, dtype=np.uint64)
        act = np.array(act, dtype=np.int32); nxt = np.array(nxt, dtype=np.uint32)
        self.nbytes = (device["num_bits"] + 7) // 8
        self.h = lib().ookd_oracle_sm_create(len(st), dur.ctypes.data, tmo.ctypes.data, off.ctypes.data,
                                             cond.ctypes.data, tdur.ctypes.data, act.ctypes.data,
                                             nxt.ctypes.data, device["num_bits"], int(sample_rate))
        if not self.h:
            raise ValueError("bad device")

    def process(self, bits):
        """-> (result, num_proc); result -1/0/1 as sm_process."""
        b = np.ascontiguousarray(bits, dtype=np.uint8)
        n = C.c_uint32(0)
        r = lib().ookd_oracle_sm_process(self.h, b.ctypes.data, len(b), C.byref(n))
        return r, n.value

    def get_state(self):
        """-> carry tuple (state, k, num_bits, prev_bit, data32) in the layout of struct ookd_sm_carry."""
        st, k, nb, pv = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        data = (C.c_uint8 * MSG_BYTES)()
        lib().ookd_oracle_sm_get_state(self.h, C.byref(st), C.byref(k), C.byref(nb), C.byref(pv), data)
        return (st.value, k.value, nb.value, pv.value, bytes(data))

    def set_state(self, carry):
        data = (C.c_uint8 * MSG_BYTES)(*carry[4])
        lib().ookd_oracle_sm_set_state(self.h, carry[0], carry[1], carry[2], carry[3], data)

    def data(self):
        p = lib().ookd_oracle_sm_data(self.h)
        return bytes(p[i] for i in range(self.nbytes))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ookd_oracle_sm_destroy(self.h)
            self.h = None


def threshold(iq_f32, thr):
    x = np.ascontiguousarray(iq_f32, dtype=np.float32).reshape(-1, 2)
    bits = np.empty(x.shape[0], dtype=np.uint8)
    lib().ookd_oracle_threshold(x.ctypes.data, x.shape[0], np.float32(thr), bits.ctypes.data)
    return bits


def rx(iq_i16, filter_stages, device, threshold_=0.1, samples_per_buffer=8192, samplerate=3000000,
       want_filtered=False, want_bits=False):
    """Whole-capture reference decode.  iq_i16: int16 array (n,2) or flat 2n.
    -> dict(n_out, n_buffers, first_bit, edges, msgs[(out_sample, buffer_idx, num_bits, bytes)], filtered, bits)"""
    iq = np.ascontiguousarray(iq_i16, dtype=np.int16).reshape(-1)
    n = iq.size // 2
    fir = Fir(filter_stages) if filter_stages else None
    dec = fir.total_decimation if fir else 1
    sm = Sm(device, samplerate // dec) if device else None
    res = _RxResult()
    rc = lib().ookd_oracle_rx(iq.ctypes.data, n, fir.h if fir else None, sm.h if sm else None,
                              np.float32(threshold_), samples_per_buffer,
                              int(want_filtered), int(want_bits), C.byref(res))
    if rc != 0:
        raise RuntimeError("ookd_oracle_rx failed")
    nbytes = (device["num_bits"] + 7) // 8 if device else 0
    out = dict(n_out=int(res.n_out), n_buffers=int(res.n_buffers), first_bit=int(res.first_bit),
               edges=np.ctypeslib.as_array(res.edges, shape=(res.n_edges,)).copy() if res.n_edges else np.zeros(0, np.uint64),
               msgs=[(int(res.msgs[i].out_sample), int(res.msgs[i].buffer_idx), int(res.msgs[i].num_bits),
                      bytes(res.msgs[i].data[:nbytes])) for i in range(res.n_msgs)],
               filtered=None, bits=None)
    if want_filtered:
        out["filtered"] = np.ctypeslib.as_array(res.filtered, shape=(res.n_out, 2)).copy()
    if want_bits:
        out["bits"] = np.ctypeslib.as_array(res.bits, shape=(res.n_out,)).copy()
    lib().ookd_oracle_rx_free(C.byref(res))
    return out


def synth(n_samples, toggles, i_on, q_on, noise_scale, seed, first_sample=0, noise_terms=4):
    tg = np.ascontiguousarray(toggles, dtype=np.uint64)
    iq = np.empty((n_samples, 2), dtype=np.int16)
    lib().ookd_oracle_synth_ex(iq.ctypes.data, first_sample, n_samples, tg.ctypes.data, len(tg),
                               int(i_on), int(q_on), int(noise_scale), int(seed), int(noise_terms))
    return iq


# Irwin-Hall(4 x U16) standard deviation in raw units; see ookd_oracle_synth.
_IH4_STD = (4.0 * (65536.0 ** 2 - 1.0) / 12.0) ** 0.5


def noise_scale_for_sigma(sigma, noise_terms=4):
    """Q24 multiplier that makes the integer noise have std `sigma` (in full-scale units, 1.0 = 2048 LSB)."""
    std = _IH4_STD * (noise_terms / 4.0) ** 0.5
    return int(round(sigma * 2048.0 / std * (1 << 24)))


def on_level(amplitude, phase_rad):
    """Integer on-level (I, Q) in SC16Q11 LSBs for a carrier of given amplitude and phase."""
    import math
    return (int(round(amplitude * 2048.0 * math.cos(phase_rad))),
            int(round(amplitude * 2048.0 * math.sin(phase_rad))))


# ----------------------------------------------------------------------------
# Transmit-side generator (capture synthesis)
# ----------------------------------------------------------------------------
def _field_width(f):
    return f["end_bit"] - f["start_bit"] + 1


def _str2u64(s):
    return int(s, 0)


def field_bits_from_str(field, s):
    """String -> raw field bits.  reference src/formatter.c:140-255 (str_to_spt)."""
    fmt = field["format"].lower()
    width = _field_width(field)
    mask = (1 << width) - 1
    scaling = np.float32(field.get("scaling", 0) or 1.0)
    offset = np.float32(field.get("offset", 0))
    if fmt in ("hex", "unsigned decimal"):
        v = int((np.float32(_str2u64(s)) - offset) / scaling)
    elif fmt == "two's complement":
        v = int((np.float32(int(s, 0)) - offset) / scaling) & mask
    elif fmt == "sign-magnitude":
        t = int(s, 0)
        v = int((np.float32(t) - offset) / scaling) & ((1 << (width - 1)) - 1)
        if t < 0:
            v |= 1 << (width - 1)
    elif fmt == "float":
        tmp = np.float32(float(s))
        v = int(np.float32((tmp - offset) / scaling)) & mask     # spt_from_float, spt.h:56-60 (truncating cast)
    elif fmt == "enumeration":
        v = None
        for e in field["enum_values"]:
            if e["string"].lower() == s.lower():
                v = _str2u64(e["value"])
                break
        if v is None:
            v = _str2u64(s)
    else:
        raise ValueError(fmt)
    if v & mask != v:
        raise ValueError(f"value too large for field {field['name']}: {s}")
    return v


def apply_field_bits(field, bits, data):
    """reference src/formatter.c:766-800 (apply_field_bits)."""
    big = field["endianness"].lower() == "big"
    src = (field["end_bit"] - field["start_bit"]) if big else 0
    for i in range(field["start_bit"], field["end_bit"] + 1):
        if bits & (1 << src):
            data[i // 8] |= 1 << (i % 8)
        else:
            data[i // 8] &= ~(1 << (i % 8)) & 0xFF
        src += -1 if big else 1


def message_bytes(device, params=None):
    """Defaults overlaid with params -> message bytes.  device.c:660-678."""
    data = bytearray((device["num_bits"] + 7) // 8)
    for f in device["fields"]:
        apply_field_bits(f, field_bits_from_str(f, f["default"]), data)
    for key, val in (params or {}).items():
        f = next(x for x in device["fields"] if x["name"].lower() == key.lower())
        apply_field_bits(f, field_bits_from_str(f, str(val)), data)
    return bytes(data)


def tx_runs(device, data, sample_rate):
    """Message bytes -> list of (level, n_samples) runs.  Restates sm_generate,
    reference src/state_machine.c:574-873 (to_sample_count :88-92)."""
    st = device["states"]
    max_bits = device["num_bits"]

    def count_of(us):
        return int(us * (float(sample_rate) / 1e6) + 0.5)

    runs = []
    g = dict(level=0, cur=0, num_bits=0)

    def append(us):
        c = count_of(us)
        if c:
            runs.append((g["level"], c))

    def pick(bit, check):
        for t in st[g["cur"]]["triggers"]:
            if check:
                a = t["action"]
                ok = (a == ACT["append_0"] and not bit) or (a == ACT["append_1"] and bit) or a == ACT["output_data"]
                if not ok:
                    continue
            c = t["cond"]
            if c == COND["msg_complete"]:
                if g["num_bits"] == max_bits:
                    return t
            elif c in (COND["always"], COND["pulse_start"], COND["pulse_end"]):
                return t
            elif c == COND["timeout"]:
                raise RuntimeError("timeout trigger reached while generating")
        return None

    def one(bit):
        done = False
        while not done:
            t = pick(bit, True) or pick(bit, False)
            if t is None:
                raise RuntimeError("no trigger")
            if st[g["cur"]]["duration_us"] == 0 and t["duration_us"] != 0:
                append(t["duration_us"])
            if t["cond"] == COND["pulse_start"]:
                g["level"] = 1
            elif t["cond"] == COND["pulse_end"]:
                g["level"] = 0
            if t["action"] in (ACT["append_0"], ACT["append_1"]):
                if g["num_bits"] < max_bits:
                    g["num_bits"] += 1
                    done = True
            elif t["action"] == ACT["output_data"]:
                done = True
            g["cur"] = t["next"]
            if st[g["cur"]]["duration_us"] != 0:
                append(st[g["cur"]]["duration_us"])

    for i in range(max_bits):
        one(bool(data[i // 8] & (1 << (i % 8))))
    one(False)
    return runs


def toggles_from_messages(device, messages, sample_rate, lead_samples, start=0):
    """Tile messages (each preceded by `lead_samples` of silence, like
    ookiedokie_tx's tx_delay, ookiedokie.c:311-337) -> (toggle positions, total length)."""
    pos = start
    tog = []
    for data in messages:
        pos += lead_samples
        level = 0
        for lvl, n in tx_runs(device, data, sample_rate):
            if lvl != level:
                tog.append(pos)
                level = lvl
            pos += n
        if level:
            tog.append(pos)
    return np.array(tog, dtype=np.uint64), pos


# ----------------------------------------------------------------------------
# Reference binaries (present only when oracle/_ref/ookiedokie was built)
# ----------------------------------------------------------------------------
def ref_binary(name="ookiedokie"):
    p = os.path.join(REF_DIR, name)
    return p if os.path.exists(p) else None


def parse_dig_csv(text):
    """--rx-rec-dig CSV (ookiedokie.c:146-169) -> (first_bit, edge positions)."""
    lines = [l for l in text.strip().splitlines() if l.strip()]
    first = int(lines[0].split(",")[1])
    edges = []
    # after the "0, b" row every transition contributes two rows: (i-1, prev) and (i, curr)
    for k in range(2, len(lines), 2):
        edges.append(int(lines[k].split(",")[0]))
    return first, np.array(edges, dtype=np.uint64)
