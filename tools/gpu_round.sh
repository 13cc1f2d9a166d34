#!/bin/bash
# Runs on the GPU box (via gpurun): parity tests, smoke, 1-GPU bench, ncu launch list + full capture of the top kernel.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench N=1"; timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
if [ "${1:-}" = "ncu" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
echo "== ncu full (raster_kernel)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:raster_kernel -s 1 -c 1 -f -o gpurun_out/prof_raster \
    python bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
fi
