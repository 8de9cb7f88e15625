# usage: tools/sweep_chunks.sh  -- state-machine chunk size x sub-windows, same box (OOKD_DEBUG adds per-warp clocks)
mkdir -p gpurun_out
for sw in 0 3; do for cb in 64 32 16; do
  bash tools/ab_bench.sh sweep_sw${sw}_cb${cb} --sub-windows $sw --chunk-buffers $cb
done; done
for cb in 64 32 16; do
  OOKD_DEBUG=1 bash tools/ab_bench.sh sweep_dbg_cb${cb} --sub-windows 0 --chunk-buffers $cb --steps 3 --warmup 2 > /dev/null
  grep "fast tail\|seed round" gpurun_out/sweep_dbg_cb${cb}.err | tail -2 | cut -c1-330
done
