"""What the ncu captures in profiles/ run: two decodes of the headline capture (C2: p3l-nexa2012 + fs32_fs4, 2^30 samples) and
two of the low-SNR capture (C3: unknown-remote1 + fs128_fs16_dec4, 2^28 samples, FMA screening).  Usage (on the GPU box):
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s <skip> -c 1 -o gpurun_out/<name> python tools/profile_target.py"""
import sys
sys.path.insert(0, ".")
import math
import numpy as np, torch
import bench
from ookiedokie_b200 import binding as B, host as H
torch.cuda.set_device(0)
which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("c2", "both"):
    n = 1 << 30
    fir = H.Fir("fs32_fs4")
    dev = H.Device("p3l-nexa2012", 3000000)
    tog, _ = bench.build_toggles(dev, n)
    d = torch.empty(n * 2, dtype=torch.int16, device="cuda")
    i_on, q_on = bench.on_level()
    B.synth(n, tog, i_on, q_on, bench.noise_scale(), bench.SEED, device_id=0, device_ptr=d.data_ptr(), noise_terms=12)
    g = B.Gpu(filter_stages=fir.stages, sm=dev.sm_spec(), threshold=0.1, samples_per_buffer=8192)
    g.want_list = False
    for _ in range(2):
        r = g.decode((d.data_ptr(), n))
    print("c2", len(r["msgs_raw"]), r["kernel_ms"])
    g.close(); del d
if which in ("c3", "both"):
    n = 1 << 28
    fir = H.Fir("fs128_fs16_dec4")
    dev = H.Device("unknown-remote1", 3000000 // 4)
    tx = H.Device("unknown-remote1", 3000000)
    msgs = [tx.message({"ID": hex(i % 256)}) for i in range(n // 200000 + 8)]
    tog, _ = tx.toggles(msgs, 12000)
    d = torch.empty(n * 2, dtype=torch.int16, device="cuda")
    B.synth(n, np.ascontiguousarray(tog), int(round(0.3 * 2048 * math.cos(0.4))), int(round(0.3 * 2048 * math.sin(0.4))),
            bench.noise_scale(0.10), 7, device_id=0, device_ptr=d.data_ptr(), noise_terms=12)
    g = B.Gpu(filter_stages=fir.stages, sm=dev.sm_spec(), threshold=0.1, samples_per_buffer=8192, flags=B.FLAG_FMA_SCREEN)
    g.want_list = False
    for _ in range(2):
        r = g.decode((d.data_ptr(), n))
    print("c3", len(r["msgs_raw"]), r["kernel_ms"], r["fir_mode"])
