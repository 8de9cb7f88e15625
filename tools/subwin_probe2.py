import sys, time, os
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from ookiedokie_b200 import binding as B, host as H
torch.cuda.set_device(0)
n = 1 << 30
fir = H.Fir(bench.FILTER_NAME)
dev = H.Device(bench.DEVICE_NAME, bench.FS // fir.total_decimation)
tog, _ = bench.build_toggles(dev, n)
i_on, q_on = bench.on_level()
d = torch.empty(n * 2, dtype=torch.int16, device="cuda")
B.synth(n, tog, i_on, q_on, bench.noise_scale(), bench.SEED, device_id=0, device_ptr=d.data_ptr(), noise_terms=bench.NOISE_TERMS)
torch.cuda.synchronize()
K = int(sys.argv[1])
m = B.MultiGpu([0] * K, filter_stages=fir.stages, sm=dev.sm_spec(), threshold=bench.THR, samples_per_buffer=bench.SPB)
halo = m.halo
ptrs = []
for s in range(K):
    sf, sn = m.shard_range(0, n, s)
    ptrs.append(d.data_ptr() + 4 * (sf - min(halo, sf)))
for _ in range(3):
    r, ex = m.decode(None, 0, n, True, device_ptrs=ptrs)
L = B.lib()
import ctypes as C
for s in range(K):
    h = C.c_void_p(L.ookd_gpu_multi_handle(m.h, s))
t0 = time.perf_counter()
r, ex = m.decode(None, 0, n, True, device_ptrs=ptrs)
print("one decode", (time.perf_counter() - t0) * 1e3, "ms; kernel_ms(max over shards)", r["kernel_ms"], "launches", r["gpu_launches"], "rounds", r["sm_rounds"])
# the same shards, one after the other on their handles (no threads): per-shard cost
from ookiedokie_b200.binding import Gpu
