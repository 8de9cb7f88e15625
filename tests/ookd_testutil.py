"""Shared helpers for the parity tests (test infrastructure; may use oracle/)."""
import math

import numpy as np

from oracle import oracle as O

FS = 3000000


def sm_spec(device, filter_stages, samplerate=FS):
    """Device dict (oracle loader) -> the dict ookiedokie_b200.binding.Gpu takes as `sm`."""
    dec = O.filter_total_decimation(filter_stages) if filter_stages else 1
    return dict(states=device["states"], num_bits=device["num_bits"], sample_rate=samplerate // dec)


def capture(device, n_msgs, sigma=0.0, amplitude=0.95, phase=0.0, seed=1, lead=12000, fields=None, tail=20000,
            samplerate=FS, glitches=()):
    """Synthetic capture: n_msgs device messages (field values vary with the index), AWGN-like
    integer noise.  -> (int16 (n,2) array, list of message bytes, toggles)."""
    msgs = []
    for i in range(n_msgs):
        params = fields(i) if fields else {}
        msgs.append(O.message_bytes(device, params))
    tog, total = O.toggles_from_messages(device, msgs, samplerate, lead)
    tog = list(tog)
    for pos, length in glitches:          # extra on-bursts in silent regions
        tog += [pos, pos + length]
    tog = np.array(sorted(tog), dtype=np.uint64)
    total += tail
    i_on, q_on = O.on_level(amplitude, phase)
    iq = O.synth(total, tog, i_on, q_on, O.noise_scale_for_sigma(sigma) if sigma > 0 else 0, seed)
    return iq, msgs, tog


def nexa_fields(i):
    return {"Channel": str(1 + i % 3), "Temperature (C)": f"{-20.0 + 0.7 * (i % 90):.1f}"}


def remote_fields(i):
    buttons = ["Power", "Pause", "P1"]
    return {"ID": hex(i % 256), "Button": buttons[i % 3]}
