"""K time shards of one device-resident 2^30-sample capture on ONE GPU, enqueued back to back from one host thread
(decode_begin x K, then decode_end x K + stitch): the in-decode sub-window pipeline, measured with the existing C ABI."""
import sys, time, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from ookiedokie_b200 import binding as B, host as H
torch.cuda.set_device(0)
L = B.lib()
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
print(f"single handle: {(time.perf_counter() - t0) * 100:.4f} ms per decode, {len(ref['msgs_raw'])} messages", flush=True)
ref_bytes = ref["msgs_raw"].tobytes()
halo = g.halo
align = int(bench.SPB * fir.total_decimation // np.gcd(bench.SPB, fir.total_decimation))

def run(K, flags, reps, sizes=None):
    hs = [B.Gpu(filter_stages=fir.stages, sm=dev.sm_spec(), threshold=bench.THR, samples_per_buffer=bench.SPB, device_id=0,
                flags=flags, sm_warmup=1) for _ in range(K)]
    for h in hs:
        h.want_list = False
    if sizes is None:
        per = (n // align + K - 1) // K * align
        cuts = [min(n, i * per) for i in range(K + 1)]
    else:
        cuts = [0]
        for s in sizes:
            cuts.append(min(n, cuts[-1] + s // align * align))
        cuts[-1] = n
    res = [B.GpuResult() for _ in range(K)]
    ex = [B.SmCarry() for _ in range(K)]
    resolves = 0
    def once():
        nonlocal resolves
        for i in range(K):
            sf, sn = cuts[i], cuts[i + 1] - cuts[i]
            p = C.c_void_p(int(d.data_ptr() + 4 * (sf - min(halo, sf))))
            rc = L.ookd_gpu_decode_begin(hs[i].h, p, 1, sf, sn, int(i == K - 1), None)
            assert rc == 0, L.ookd_gpu_last_error(hs[i].h)
        for i in range(K):
            rc = L.ookd_gpu_decode_end(hs[i].h, C.byref(ex[i]), C.byref(res[i]))
            assert rc == 0, L.ookd_gpu_last_error(hs[i].h)
            if i > 0 and bytes(res[i].entry_used) != bytes(ex[i - 1]):
                rc = L.ookd_gpu_resolve(hs[i].h, C.byref(ex[i - 1]), C.byref(ex[i]), C.byref(res[i]))
                assert rc == 0, L.ookd_gpu_last_error(hs[i].h)
                resolves += 1
    for _ in range(3):
        once()
    resolves = 0
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    dt = (time.perf_counter() - t0) / reps * 1e3
    msgs = b"".join(hs[i]._result(res[i])["msgs_raw"].tobytes() for i in range(K))
    print(f"K {K} flags {flags} sizes {sizes}: {dt:.4f} ms per decode, resolves/decode {resolves / reps:.2f}, identical {msgs == ref_bytes}, "
          f"per-shard kernel spans {[round(float(r.kernel_ms), 3) for r in res]}", flush=True)
    for h in hs:
        h.close()

for flags in (0, B.FLAG_SHARE_SMS):
    for K in (2, 4, 8):
        run(K, flags, 10)
run(2, 0, 10, sizes=[(n * 7) // 8, n // 8])
run(3, 0, 10, sizes=[n // 2, (n * 3) // 8, n // 8])
run(2, B.FLAG_SHARE_SMS, 10, sizes=[(n * 7) // 8, n // 8])
