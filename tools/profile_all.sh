#!/bin/bash
# One GPU-box call that produces the round's evidence files under gpurun_out/ (copy the ones to keep into profiles/):
#   r2_bench_n1.json / r2_bench_reference.json   both bench arms, default arguments
#   r2_launches_bench_f64.csv                    ncu launch list of a short bench run (per-launch durations)
#   r2_ncu_rb_stream_<dtype>_<depth>_<noise>_<src>.txt   ncu --set full summary (tools/ncu_summary.py) of one launch per variant
# The .ncu-rep files are summarised on the box and deleted (gpurun returns at most 64 MiB).
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>/dev/null
B="python bench.py --steps 2 --warmup 1 --no-variants --no-e2e --no-cpu-baseline"
$B > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench_f64.csv $B > gpurun_out/ncu_l.log 2>&1
for cfg in "f64 4 0 1" "f32 4 0 1" "f64 4 1 1" "f32 4 1 1" "f64 1 0 1"; do
  tag=$(echo $cfg | tr ' ' '_')
  python tools/one_config.py $cfg > gpurun_out/plain_$tag.log 2>&1 && ncu --set full --clock-control none -k regex:rb_stream -s 2 -c 1 -f -o /tmp/prof_$tag python tools/one_config.py $cfg > gpurun_out/ncu_$tag.log 2>&1
  python tools/ncu_summary.py /tmp/prof_$tag.ncu-rep > gpurun_out/r2_ncu_rb_stream_$tag.txt 2>&1
  tail -1 gpurun_out/plain_$tag.log
done
