"""Where does a batch of short captures spend its time: host enqueue (decode_begin) or waiting for the GPU (decode_end)?"""
import sys, time, ctypes as C
sys.path.insert(0, ".")
import numpy as np, torch
import bench
from ookiedokie_b200 import binding as B, host as H
torch.cuda.set_device(0)
L = B.lib()
n = 1 << 24
fir = H.Fir("fs64_fs8")
devs = [H.Device(nm, bench.FS // fir.total_decimation) for nm in ("p3l-nexa2012", "unknown-remote1")]
bufs = []
for i in range(64):
    kind = i % 2
    msgs = [devs[kind].message({}) for _ in range(n // (180000 if kind else 400000) + 2)]
    tog, total = devs[kind].toggles(msgs, bench.LEAD)
    sigma = [0.0, 0.02, 0.05][i % 3]
    d = torch.empty(n * 2, dtype=torch.int16, device="cuda")
    B.synth(n, np.ascontiguousarray(tog), 1488, 1253, bench.noise_scale(sigma) if sigma else 0, 1000 + i, device_id=0,
            device_ptr=d.data_ptr(), noise_terms=bench.NOISE_TERMS)
    bufs.append(d)
torch.cuda.synchronize()
for per in (1, 4, 16):
    gpus = [B.Gpu(filter_stages=fir.stages, sm=devs[k].sm_spec(), threshold=bench.THR, samples_per_buffer=bench.SPB, device_id=0)
            for k in range(2) for _ in range(per)]
    res, ex = B.GpuResult(), B.SmCarry()
    def run():
        tb = te = 0.0
        inflight = [None] * len(gpus)
        for j, d in enumerate(bufs):
            hi = (j % 2) * per + (j // 2) % per
            if inflight[hi] is not None:
                t = time.perf_counter()
                assert L.ookd_gpu_decode_end(gpus[hi].h, C.byref(ex), C.byref(res)) == 0
                te += time.perf_counter() - t
            t = time.perf_counter()
            assert L.ookd_gpu_decode_begin(gpus[hi].h, C.c_void_p(d.data_ptr()), 1, 0, n, 1, None) == 0
            tb += time.perf_counter() - t
            inflight[hi] = j
        for hi in range(len(gpus)):
            if inflight[hi] is not None:
                t = time.perf_counter()
                assert L.ookd_gpu_decode_end(gpus[hi].h, C.byref(ex), C.byref(res)) == 0
                te += time.perf_counter() - t
        return tb, te
    run(); run()
    t0 = time.perf_counter()
    tb, te = run()
    dt = time.perf_counter() - t0
    print(f"{2 * per} handles: {dt * 1e3:.2f} ms for {len(bufs)} captures = {dt / len(bufs) * 1e6:.0f} us each; in decode_begin {tb / len(bufs) * 1e6:.0f} us, "
          f"in decode_end {te / len(bufs) * 1e6:.0f} us per capture", flush=True)
    for g in gpus:
        g.close()
