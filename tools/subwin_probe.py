"""One 2^30-sample device-resident capture decoded as K time shards on ONE GPU (ookd_gpu_multi with K handles on device 0):
how much of the tail hides behind the next shard's screening kernel?"""
import sys, time
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
g = B.Gpu(filter_stages=fir.stages, sm=dev.sm_spec(), threshold=bench.THR, samples_per_buffer=bench.SPB, device_id=0)
g.want_list = False
for _ in range(3):
    ref = g.decode((d.data_ptr(), n))
t0 = time.perf_counter()
for _ in range(10):
    ref = g.decode((d.data_ptr(), n))
print(f"single handle: {(time.perf_counter() - t0) * 100:.4f} ms per decode, {len(ref['msgs_raw'])} messages")
ref_bytes = ref["msgs_raw"].tobytes()
import os
print('env', {k: v for k, v in os.environ.items() if k.startswith('OOKD_')}, flush=True)
for flags in (0,):
    for K in (2, 4, 8, 16):
        m = B.MultiGpu([0] * K, filter_stages=fir.stages, sm=dev.sm_spec(), threshold=bench.THR, samples_per_buffer=bench.SPB, flags=flags)
        halo = m.halo
        ptrs = []
        for s in range(K):
            sf, sn = m.shard_range(0, n, s)
            ha = min(halo, sf)
            ptrs.append(d.data_ptr() + 4 * (sf - ha))
        for _ in range(3):
            r, ex = m.decode(None, 0, n, True, device_ptrs=ptrs)
        import ctypes as C
        L = B.lib()
        arr = (C.c_void_p * K)(*[int(p) for p in ptrs])
        res, exc = B.GpuResult(), B.SmCarry()
        t0 = time.perf_counter()
        for _ in range(10):                                   # (the C call alone: no Python-side message conversion)
            rc = L.ookd_gpu_multi_decode(m.h, C.cast(arr, C.c_void_p), 1, 0, n, 1, None, C.byref(exc), C.byref(res))
            assert rc == 0
        dt = (time.perf_counter() - t0) * 100
        same = r["msgs_raw"].tobytes() == ref_bytes
        print(f"K {K} flags {flags}: {dt:.4f} ms per decode, {len(r['msgs_raw'])} messages, identical {same}, rounds {r['sm_rounds']}")
        m.close()
