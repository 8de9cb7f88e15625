"""Per-noise-level cost of a 2^24-sample capture through fs64_fs8 (what bounds the mixed batch)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from ookiedokie_b200 import binding as B, host as H
torch.cuda.set_device(0)
n = 1 << 24
fir = H.Fir("fs64_fs8")
for name in ("p3l-nexa2012", "unknown-remote1"):
    dev = H.Device(name, 3000000)
    msgs = [dev.message({}) for _ in range(n // (180000 if "remote" in name else 400000) + 2)]
    tog, total = dev.toggles(msgs, 12000)
    for sigma in (0.0, 0.02, 0.05):
        d = torch.empty(n * 2, dtype=torch.int16, device="cuda")
        B.synth(n, np.ascontiguousarray(tog), 1488, 1253, bench.noise_scale(sigma) if sigma else 0, 7, device_id=0, device_ptr=d.data_ptr(), noise_terms=12)
        torch.cuda.synchronize()
        g = B.Gpu(filter_stages=fir.stages, sm=dev.sm_spec(), threshold=0.1, samples_per_buffer=8192)
        g.want_list = False
        for _ in range(3):
            r = g.decode((d.data_ptr(), n))
        t0 = time.perf_counter()
        for i in range(16):
            r = g.decode((d.data_ptr(), n))
        dt = (time.perf_counter() - t0) / 16
        print(f"{name} sigma {sigma}: {dt * 1e3:.3f} ms per decode, kernel span {r['kernel_ms']:.3f}, fir {r['fir_ms']:.3f}, screen {r['screen_ms']:.3f}, mode {r['fir_mode']}, "
              f"edges {r['n_edges']}, msgs {len(r['msgs_raw'])}, refined {r['refined_blocks']}, rounds {r['sm_rounds']}, syncs {r['host_syncs']}, launches {r['gpu_launches']}")
        g.close()
