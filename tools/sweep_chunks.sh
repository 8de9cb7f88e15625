mkdir -p gpurun_out
for cb in 8 16 32 64; do for fl in 0 1024; do
  OOKD_DEBUG=1 timeout 120 python bench.py --steps 15 --warmup 5 --no-e2e --no-cpu --no-configs --pipelined-depth 0 --chunk-buffers $cb --flags $fl 2> gpurun_out/r3o_${cb}_${fl}.err | grep "^{" > gpurun_out/r3o_${cb}_${fl}.json
  python - <<P
import json
d=json.loads(open("gpurun_out/r3o_${cb}_${fl}.json").read().strip().splitlines()[-1])
print("cb", $cb, "flags", $fl, "ms", round(d["ms_per_step"],4), "lat", round(d.get("step_latency_ms",0),4))
P
  grep "fast tail\|seed round" gpurun_out/r3o_${cb}_${fl}.err | tail -2 | cut -c1-200
done; done
