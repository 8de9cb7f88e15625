"""ctypes view of the C host front end (ookiedokie_b200/host -> lib/libookd_host.so).

Glue for tests and bench.py: the loaders, formatter and transmit-side generator it exposes are the C
ones the `ookiedokie-b200` CLI uses (ookd_host.h); nothing is re-implemented here.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import binding as B

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libookd_host.so")
CLI_PATH = os.path.join(_HERE, "bin", "ookiedokie-b200")
DATA_DIR = os.path.join(_HERE, "data")

EXPORTS = [
    "ookd_log_set_verbosity", "ookd_log", "ookd_find_device_file", "ookd_find_filter_file", "ookd_data_dir",
    "ookd_keyval_list_init", "ookd_keyval_list_append", "ookd_keyval_list_clear", "ookd_keyval_list_deinit",
    "ookd_fir_init", "ookd_fir_deinit", "ookd_fir_get_total_decimation", "ookd_fir_desc",
    "ookd_device_init", "ookd_device_deinit", "ookd_device_sm_desc", "ookd_device_num_bits", "ookd_device_name",
    "ookd_device_format", "ookd_device_message", "ookd_device_generate_runs", "ookd_device_toggles",
    "ookd_cfg_init", "ookd_rx", "ookd_tx", "ookd_rx_print", "ookd_rx_request_stop",
]


class KeyVal(C.Structure):
    _fields_ = [("key", C.c_char_p), ("value", C.c_char_p)]


class KeyValList(C.Structure):
    _fields_ = [("items", C.POINTER(KeyVal)), ("n", C.c_size_t), ("cap", C.c_size_t)]


def build_library():
    subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "csrc")], check=True)
    subprocess.run(["make", "-s", "-C", os.path.join(_HERE, "host")], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        B.lib()                                     # libookd_gpu.so first (rpath $ORIGIN also finds it)
        if not os.path.exists(LIB_PATH):
            build_library()
        os.environ.setdefault("OOKD_DATA_DIR", DATA_DIR + "/")
        L = C.CDLL(LIB_PATH)
        L.ookd_data_dir.restype = C.c_char_p
        L.ookd_fir_init.restype = C.c_void_p
        L.ookd_fir_init.argtypes = [C.c_char_p]
        L.ookd_fir_deinit.argtypes = [C.c_void_p]
        L.ookd_fir_get_total_decimation.restype = C.c_uint
        L.ookd_fir_get_total_decimation.argtypes = [C.c_void_p]
        L.ookd_fir_desc.restype = C.POINTER(B.FilterDesc)
        L.ookd_fir_desc.argtypes = [C.c_void_p]
        L.ookd_device_init.restype = C.c_void_p
        L.ookd_device_init.argtypes = [C.c_char_p, C.c_uint]
        L.ookd_device_deinit.argtypes = [C.c_void_p]
        L.ookd_device_sm_desc.restype = C.POINTER(B.SmDesc)
        L.ookd_device_sm_desc.argtypes = [C.c_void_p]
        L.ookd_device_num_bits.restype = C.c_uint
        L.ookd_device_num_bits.argtypes = [C.c_void_p]
        L.ookd_device_name.restype = C.c_char_p
        L.ookd_device_name.argtypes = [C.c_void_p]
        L.ookd_keyval_list_init.argtypes = [C.POINTER(KeyValList)]
        L.ookd_keyval_list_append.restype = C.c_bool
        L.ookd_keyval_list_append.argtypes = [C.POINTER(KeyValList), C.c_char_p, C.c_char_p]
        L.ookd_keyval_list_clear.argtypes = [C.POINTER(KeyValList)]
        L.ookd_keyval_list_deinit.argtypes = [C.POINTER(KeyValList)]
        L.ookd_device_format.restype = C.c_bool
        L.ookd_device_format.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(KeyValList)]
        L.ookd_device_message.restype = C.c_bool
        L.ookd_device_message.argtypes = [C.c_void_p, C.POINTER(KeyValList), C.c_void_p]
        L.ookd_device_generate_runs.restype = C.c_bool
        L.ookd_device_generate_runs.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.POINTER(C.c_uint32)),
                                                C.POINTER(C.c_size_t)]
        L.ookd_device_toggles.restype = C.c_bool
        L.ookd_device_toggles.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint64,
                                          C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_size_t),
                                          C.POINTER(C.c_uint64)]
        L.ookd_log_set_verbosity.argtypes = [C.c_int]
        _lib = L
    return _lib


_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]


class Fir:
    """fir_init: filter JSON -> stage list."""

    def __init__(self, name):
        self.h = lib().ookd_fir_init(name.encode())
        if not self.h:
            raise ValueError(f"cannot load filter {name}")
        d = lib().ookd_fir_desc(self.h).contents
        self.total_decimation = int(lib().ookd_fir_get_total_decimation(self.h))
        self.stages = [(int(d.decimation[i]),
                        np.ctypeslib.as_array(d.taps[i], shape=(d.num_taps[i],)).astype(np.float32).copy())
                       for i in range(d.num_stages)]

    def __del__(self):
        # at interpreter shutdown module globals may already be gone: then the process is exiting anyway
        if getattr(self, "h", None) and lib is not None:
            lib().ookd_fir_deinit(self.h)
            self.h = None


class Device:
    """device_init: device JSON -> state machine description, formatter and TX generator."""

    def __init__(self, name, sample_rate):
        self.h = lib().ookd_device_init(name.encode(), int(sample_rate))
        if not self.h:
            raise ValueError(f"cannot load device {name}")
        self.num_bits = int(lib().ookd_device_num_bits(self.h))
        self.nbytes = (self.num_bits + 7) // 8
        self.sample_rate = int(sample_rate)
        self.name = lib().ookd_device_name(self.h).decode()

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            lib().ookd_device_deinit(self.h)
            self.h = None

    def sm_spec(self):
        """The dict ookiedokie_b200.binding.Gpu takes as `sm` (copied out of struct ookd_sm_desc)."""
        d = lib().ookd_device_sm_desc(self.h).contents
        states = []
        for i in range(d.num_states):
            s = d.states[i]
            trigs = [dict(cond=int(d.triggers[q].cond), action=int(d.triggers[q].action),
                          next=int(d.triggers[q].next_state), duration_us=int(d.triggers[q].duration_us))
                     for q in range(s.first_trigger, s.first_trigger + s.num_triggers)]
            states.append(dict(duration_us=int(s.duration_us), timeout_us=int(s.timeout_us), triggers=trigs))
        return dict(states=states, num_bits=int(d.max_bits), sample_rate=int(d.sample_rate))

    def format(self, data):
        """Message bytes -> [(key, value)] (timestamp entry first when the device has a ts_mode)."""
        kv = KeyValList()
        lib().ookd_keyval_list_init(C.byref(kv))
        buf = (C.c_uint8 * B.MSG_BYTES)(*bytes(data))
        ok = lib().ookd_device_format(self.h, buf, C.byref(kv))
        out = [(kv.items[i].key.decode(), kv.items[i].value.decode()) for i in range(kv.n)]
        lib().ookd_keyval_list_deinit(C.byref(kv))
        if not ok:
            raise RuntimeError("ookd_device_format failed")
        return out

    def message(self, params=None):
        """Defaults overlaid with params -> message bytes."""
        kv = KeyValList()
        lib().ookd_keyval_list_init(C.byref(kv))
        for k, v in (params or {}).items():
            lib().ookd_keyval_list_append(C.byref(kv), str(k).encode(), str(v).encode())
        buf = (C.c_uint8 * B.MSG_BYTES)()
        ok = lib().ookd_device_message(self.h, C.byref(kv), buf)
        lib().ookd_keyval_list_deinit(C.byref(kv))
        if not ok:
            raise ValueError("bad device parameters")
        return bytes(buf[:self.nbytes])

    def runs(self, data):
        p = C.POINTER(C.c_uint32)()
        n = C.c_size_t()
        buf = (C.c_uint8 * B.MSG_BYTES)(*bytes(data))
        if not lib().ookd_device_generate_runs(self.h, buf, C.byref(p), C.byref(n)):
            raise RuntimeError("ookd_device_generate_runs failed")
        out = [(int(p[2 * i]), int(p[2 * i + 1])) for i in range(n.value)]
        _libc.free(p)
        return out

    def toggles(self, messages, lead_samples, start=0):
        """Tile messages -> (uint64 toggle positions, total samples)."""
        blob = b"".join(bytes(m) for m in messages)
        p = C.POINTER(C.c_uint64)()
        n = C.c_size_t()
        total = C.c_uint64()
        arr = (C.c_uint8 * max(len(blob), 1)).from_buffer_copy(blob or b"\0")
        if not lib().ookd_device_toggles(self.h, arr, len(messages), int(lead_samples), int(start), C.byref(p),
                                         C.byref(n), C.byref(total)):
            raise RuntimeError("ookd_device_toggles failed")
        out = np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint64)
        _libc.free(p)
        return out, int(total.value)


def run_cli(args, **kw):
    """Run the ookiedokie-b200 binary; returns CompletedProcess."""
    env = dict(os.environ)
    env.setdefault("OOKD_DATA_DIR", DATA_DIR + "/")
    return subprocess.run([CLI_PATH] + list(args), capture_output=True, text=True, env=env, **kw)
