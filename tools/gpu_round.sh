#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests, smoke, 1-GPU bench, A/B variants, ncu launch list + full captures.
# Every step is bounded by its own timeout so that a hung kernel can never hold the box until gpurun's limit.
#   VARIANTS="a b"   tools/dev/variants/<name>.so builds to bench    E2E="4 16 d"  extra e2e runs (chunk counts, d = direct)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench N=1"; timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
for e in ${E2E:-}; do
  if [ "$e" = "d" ]; then fl="--e2e-direct 1"; else fl="--e2e-chunks $e"; fi
  echo "== e2e $fl"; timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-auto $fl > gpurun_out/bench_e2e_$e.json 2> gpurun_out/bench_e2e_$e.err; tail -1 gpurun_out/bench_e2e_$e.err
  python -c "import json;d=json.load(open('gpurun_out/bench_e2e_$e.json'));print('value',round(d['value']),'e2e',round(d['e2e']['value']),d['e2e']['ms_per_step'])"
done
for v in ${VARIANTS:-}; do
  echo "== variant $v"; timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-auto --lib tools/dev/variants/$v.so > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; tail -1 gpurun_out/bench_$v.err
  python -c "import json;d=json.load(open('gpurun_out/bench_$v.json'));print('value',round(d['value']),d['stage_ms'],'e2e',round(d['e2e']['value']))"
done
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
if [ "${1:-}" = "ncu" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-300
echo "== ncu full (raster_kernel, line_cover_kernel, build_geometry_kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"raster_kernel|line_cover_kernel|build_geometry_kernel" -s 3 -c 3 -f -o gpurun_out/prof_r01 \
    python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-auto > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
echo "== ncu launch list of the f3 / f4 kernels"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"auto_|png_|class_reach|entity_box" -c 60 --csv --log-file gpurun_out/launches_f3f4.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/ncu_f3f4.log 2>&1
tail -1 gpurun_out/ncu_f3f4.log | cut -c1-200
echo "== bench C3 (@2x)"
timeout 600 python bench.py --workload C3 --steps 5 --warmup 3 --skip-auto > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -1 gpurun_out/bench_c3.err
python -c "import json;d=json.load(open('gpurun_out/bench_c3.json'));print('C3 value',round(d['value']),'e2e',round(d['e2e']['value']),'cpu',d['cpu_baseline'],'diff',d['max_abs_diff_rgb_vs_cpu'])"
ls -la gpurun_out
fi
