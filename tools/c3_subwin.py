import sys
sys.path.insert(0, ".")
import bench, torch
torch.cuda.set_device(0)
for k in (0, 2, 3, 4, 6):
    r = bench.measure_c3(0, 6548.2, sub_windows=k)
    print("c3 K", k, round(r["ms_per_step"], 4), "ms", round(r["value"]), "Ms/s fir", round(r["fir_kernel_ms"], 4), r["messages_decoded"], r["fir_mode"], flush=True)
