"""Batch-decode diagnostics: latency of one 2^24-sample decode alone, and throughput with several in flight."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from ookiedokie_b200 import binding as B, host as H
torch.cuda.set_device(0)
n = 1 << 24
fir = H.Fir("fs64_fs8")
dev = H.Device("p3l-nexa2012", 3000000)
msgs = [dev.message({}) for _ in range(n // 400000 + 2)]
tog, total = dev.toggles(msgs, 12000)
bufs = []
for i in range(16):
    d = torch.empty(n * 2, dtype=torch.int16, device="cuda")
    B.synth(n, np.ascontiguousarray(tog), 1488, 1253, bench.noise_scale(0.02), 1000 + i, device_id=0, device_ptr=d.data_ptr(), noise_terms=12)
    bufs.append(d)
torch.cuda.synchronize()
for flags in (0, B.FLAG_NO_GRAPH, B.FLAG_SYNC_TAIL):
    g = B.Gpu(filter_stages=fir.stages, sm=dev.sm_spec(), threshold=0.1, samples_per_buffer=8192, flags=flags)
    g.want_list = False
    for _ in range(3):
        r = g.decode((bufs[0].data_ptr(), n))
    t0 = time.perf_counter()
    for i in range(32):
        r = g.decode((bufs[i % 16].data_ptr(), n))
    dt = (time.perf_counter() - t0) / 32
    print(f"flags {flags}: one handle, sequential: {dt * 1e3:.3f} ms per decode, kernel span {r['kernel_ms']:.3f} ms, fir {r['fir_ms']:.3f}, "
          f"launches {r['gpu_launches']}, syncs {r['host_syncs']}, msgs {len(r['msgs_raw'])}")
    # k handles, begin all then end all
    for k in (2, 4, 8, 16):
        hs = [B.Gpu(filter_stages=fir.stages, sm=dev.sm_spec(), threshold=0.1, samples_per_buffer=8192, flags=flags) for _ in range(k)]
        for h in hs:
            h.want_list = False
            h.decode((bufs[0].data_ptr(), n))
            h.decode((bufs[1].data_ptr(), n))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 4
        for rep in range(reps):
            for j, h in enumerate(hs):
                h.decode_begin((bufs[j % 16].data_ptr(), n), 0, n, True)
            for h in hs:
                h.decode_end()
        dt = (time.perf_counter() - t0) / (reps * k)
        print(f"   {k} handles begin-all / end-all from one thread: {dt * 1e3:.3f} ms per decode")
        for h in hs:
            h.close()
    g.close()
