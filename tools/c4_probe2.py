"""Timing of a WARM 2^33-sample shard (second shard of a longer capture, entered from warm-up history)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from ookiedokie_b200 import binding as B, host as H
torch.cuda.set_device(0)
fir = H.Fir("fs32_fs4")
dev = H.Device("p3l-nexa2012", 3000000)
for log2n in (30, 33):
    n = 1 << log2n
    tog, _ = bench.build_toggles(dev, 2 * n)
    g = B.Gpu(filter_stages=fir.stages, sm=dev.sm_spec(), threshold=0.1, samples_per_buffer=8192, sm_warmup=1)
    g.want_list = False
    halo = g.halo
    d = torch.empty((halo + n) * 2, dtype=torch.int16, device="cuda")
    B.synth(halo + n, tog, 1488, 1253, bench.noise_scale(), bench.SEED, first_sample=n - halo, device_id=0, device_ptr=d.data_ptr(), noise_terms=12)
    torch.cuda.synchronize()
    for rep in range(4):
        t0 = time.perf_counter()
        r, ex = g.decode_shard((d.data_ptr(), halo + n), n, n, True, None)
        dt = time.perf_counter() - t0
        print(f"2^{log2n} warm rep {rep}: {dt * 1e3:.3f} ms, kernel span {r['kernel_ms']:.3f}, fir {r['fir_ms']:.3f}, screen {r['screen_ms']:.3f}, "
              f"edges {r['n_edges']}, msgs {len(r['msgs_raw'])}, rounds {r['sm_rounds']}, syncs {r['host_syncs']}, launches {r['gpu_launches']}, prov {r['entry_is_provisional']}")
    g.close()
    del d
    torch.cuda.empty_cache()
