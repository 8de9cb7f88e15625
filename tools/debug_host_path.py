"""Debug aid: repeated device-input and host-input decodes of the bench capture must agree."""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as BN
import torch
from ookiedokie_b200 import binding as B, host as H

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0
fir = H.Fir(BN.FILTER_NAME)
dev = H.Device(BN.DEVICE_NAME, BN.FS)
gpu = B.Gpu(filter_stages=fir.stages, sm=dev.sm_spec(), threshold=BN.THR, samples_per_buffer=BN.SPB, device_id=0, flags=flags)
gpu.want_list = False
tog, _ = BN.build_toggles(dev, n)
i_on, q_on = BN.on_level()
d_iq = torch.empty((n * 2,), dtype=torch.int16, device="cuda")
B.synth(n, tog, i_on, q_on, BN.noise_scale(), BN.SEED, first_sample=0, device_id=0, device_ptr=d_iq.data_ptr())
torch.cuda.synchronize()
hptr = B.lib().ookd_gpu_host_alloc(n * 4)
assert B.lib().ookd_gpu_memcpy_d2h(0, hptr, d_iq.data_ptr(), n * 4) == 0
h_iq = np.ctypeslib.as_array(ctypes.cast(hptr, ctypes.POINTER(ctypes.c_int16)), shape=(n * 2,))

ref = None
for kind in ["dev"] * reps + ["host"] * reps + ["dev"] * 2:
    arg = (d_iq.data_ptr(), n) if kind == "dev" else h_iq
    res, ex = gpu.decode_shard(arg, 0, n, True, None)
    fb, edges = gpu.edges()
    cur = (res["msgs_raw"].copy(), edges.copy(), fb)
    if ref is None:
        ref = cur
        print(kind, "ref: msgs", len(cur[0]), "edges", len(cur[1]), "refined", res["refined_blocks"], flush=True)
        continue
    same_m = np.array_equal(cur[0], ref[0])
    same_e = len(cur[1]) == len(ref[1]) and np.array_equal(cur[1], ref[1])
    print(kind, "msgs equal", same_m, "edges equal", same_e, "n_msgs", len(cur[0]), "n_edges", len(cur[1]),
          "refined", res["refined_blocks"], "rounds", res["sm_rounds"], flush=True)
    if not same_e:
        a, b = ref[1], cur[1]
        m = min(len(a), len(b))
        d = np.nonzero(a[:m] != b[:m])[0]
        i = int(d[0]) if len(d) else m
        print("  first differing edge idx", i, "ref", a[max(0, i - 2):i + 3], "cur", b[max(0, i - 2):i + 3],
              "tile", int(a[i] if i < len(a) else b[i]) // 4096, "piece", int(a[i] if i < len(a) else b[i]) // (16 << 20))
    elif not same_m:
        a, b = ref[0], cur[0]
        m = min(len(a), len(b))
        d = np.nonzero(a[:m] != b[:m])[0]
        i = int(d[0]) if len(d) else m
        print("  first differing msg idx", i, a[max(0, i - 1):i + 2], b[max(0, i - 1):i + 2])
