"""Timing of one 2^33-sample shard (BASELINE configs[3] per-GPU size) against 2^30."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from ookiedokie_b200 import binding as B, host as H
torch.cuda.set_device(0)
fir = H.Fir("fs32_fs4")
dev = H.Device("p3l-nexa2012", 3000000)
for log2n in (30, 32, 33):
    n = 1 << log2n
    tog, _ = bench.build_toggles(dev, n)
    d = torch.empty(n * 2, dtype=torch.int16, device="cuda")
    B.synth(n, tog, 1488, 1253, bench.noise_scale(), bench.SEED, device_id=0, device_ptr=d.data_ptr(), noise_terms=12)
    torch.cuda.synchronize()
    g = B.Gpu(filter_stages=fir.stages, sm=dev.sm_spec(), threshold=0.1, samples_per_buffer=8192)
    g.want_list = False
    for _ in range(2):
        r = g.decode((d.data_ptr(), n))
    t0 = time.perf_counter()
    for _ in range(3):
        r = g.decode((d.data_ptr(), n))
    dt = (time.perf_counter() - t0) / 3
    print(f"2^{log2n}: {dt * 1e3:.3f} ms per decode ({4 * n / dt / 1e9:.0f} GB/s), kernel span {r['kernel_ms']:.3f}, fir {r['fir_ms']:.3f}, screen {r['screen_ms']:.3f}, "
          f"edges {r['n_edges']}, msgs {len(r['msgs_raw'])}, rounds {r['sm_rounds']}, syncs {r['host_syncs']}, launches {r['gpu_launches']}, refined {r['refined_blocks']}")
    g.close()
    del d
    torch.cuda.empty_cache()
