# usage: tools/ab_bench.sh <tag> [extra bench args]   -- short device-resident bench, prints ms/step and the device span
tag=$1; shift
timeout 200 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-configs --pipelined-depth 0 "$@" 2> gpurun_out/${tag}.err | grep "^{" > gpurun_out/${tag}.json
python - <<P
import json
d=json.loads(open("gpurun_out/${tag}.json").read().strip().splitlines()[-1])
print("${tag}", "ms", round(d["ms_per_step"],4), "lat", round(d.get("step_latency_ms",0),4), "fir", round(d.get("fir_stage_ms_per_step",0),4), "job_frac", d["roofline"].get("job_frac"))
P
