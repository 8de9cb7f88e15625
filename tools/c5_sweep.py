import json, sys
sys.path.insert(0, ".")
import bench, torch
torch.cuda.set_device(0)
for per in (4, 8, 16, 32):
    r = bench.measure_c5(0, 1, 0, 6548.2, per=per)
    print(per, round(r["ms_total"], 2), round(r["value"]), round(r["hbm_frac_job"], 4), r["messages_decoded"], r["launches_per_capture"])
