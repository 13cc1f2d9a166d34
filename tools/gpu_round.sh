#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests, smoke, 1-GPU bench, A/B variants, ncu launch list + full captures.
# Every step is bounded by its own timeout so that a hung kernel can never hold the box until gpurun's limit.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench N=1"; timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
echo "== bench staged e2e"; timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --e2e-direct 0 > gpurun_out/bench_staged.json 2> gpurun_out/bench_staged.err; cat gpurun_out/bench_staged.json
for v in ${VARIANTS:-}; do
  echo "== variant $v"; timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --lib tools/dev/variants/$v.so > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; tail -1 gpurun_out/bench_$v.err; cat gpurun_out/bench_$v.json
done
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
if [ "${1:-}" = "ncu" ]; then
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
echo "== ncu full (raster_kernel, line_cover_kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"raster_kernel|line_cover_kernel" -s 2 -c 2 -f -o gpurun_out/prof_r01 \
    python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
fi
