"""A warm 2^33-sample shard (32 GiB, second shard of a longer capture) decoded with K sub-windows: which K suits it?"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from ookiedokie_b200 import binding as B, host as H
torch.cuda.set_device(0)
fir = H.Fir("fs32_fs4")
dev = H.Device("p3l-nexa2012", 3000000)
n = 1 << 33
tog, _ = bench.build_toggles(dev, 2 * n)
g0 = B.Gpu(filter_stages=fir.stages, sm=dev.sm_spec(), threshold=0.1, samples_per_buffer=8192, sm_warmup=1)
halo = g0.halo
g0.close()
d = torch.empty((halo + n) * 2, dtype=torch.int16, device="cuda")
B.synth(halo + n, tog, 1488, 1253, bench.noise_scale(), bench.SEED, first_sample=n - halo, device_id=0, device_ptr=d.data_ptr(), noise_terms=12)
torch.cuda.synchronize()
ref = None
for K in (0, 3, 6, 12, 24):
    g = B.Gpu(filter_stages=fir.stages, sm=dev.sm_spec(), threshold=0.1, samples_per_buffer=8192, sm_warmup=1, sub_windows=K)
    g.want_list = False
    for rep in range(3):
        r, ex = g.decode_shard((d.data_ptr(), halo + n), n, n, True, None)
    t0 = time.perf_counter()
    for rep in range(5):
        r, ex = g.decode_shard((d.data_ptr(), halo + n), n, n, True, None)
    dt = (time.perf_counter() - t0) / 5
    b = r["msgs_raw"].tobytes()
    if ref is None:
        ref = (b, ex)
    print(f"K {K}: {dt * 1e3:.3f} ms per decode ({4 * n / dt / 1e9:.0f} GB/s), span {r['kernel_ms']:.3f}, screen {r['screen_ms']:.3f}, msgs {len(r['msgs_raw'])}, "
          f"rounds {r['sm_rounds']}, launches {r['gpu_launches']}, same {b == ref[0] and ex == ref[1]}", flush=True)
    g.close()
