#!/bin/bash
# One gpurun call of the development loop; everything lands in gpurun_out/<tag>_*.
# Usage: tools/gpu_session.sh <tag> [smoke tests bench benchref launches ncufull cfg5 sweep ...]
set -u
TAG=${1:-s}
shift || true
WHAT=${*:-"smoke tests bench benchref launches ncufull"}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on -c 1 -f"
QUIET="--skip-e2e --skip-cpu --others ''"
for w in $WHAT; do
  case $w in
    smoke) timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" ;;
    tests) timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${TAG}_tests.log ;;
    bench) timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/${TAG}_bench.json ;;
    benchref) timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_benchref.json 2> gpurun_out/${TAG}_benchref.err; echo "benchref rc=$?"; cat gpurun_out/${TAG}_benchref.json ;;
    cfg5) timeout 900 python bench.py --workload cfg5 --steps 3 --skip-e2e --skip-cpu --others '' > gpurun_out/${TAG}_bench_cfg5.json 2> gpurun_out/${TAG}_bench_cfg5.err; echo "cfg5 rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench_cfg5.json; tail -3 gpurun_out/${TAG}_bench_cfg5.err ;;
    cfg1) timeout 300 python bench.py --workload cfg1 --steps 20 --others '' > gpurun_out/${TAG}_bench_cfg1.json 2> gpurun_out/${TAG}_bench_cfg1.err; echo "cfg1 rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench_cfg1.json ;;
    sweep) timeout 900 python tools/sweep.py > gpurun_out/${TAG}_sweep.jsonl 2> gpurun_out/${TAG}_sweep.err; echo "sweep rc=$?"; tail -3 gpurun_out/${TAG}_sweep.err ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
        python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu-launches rc=$?" ;;
    ncufull)
      timeout 600 $NCU -k 'regex:^k_spmm$' -s 3 -o gpurun_out/${TAG}_prof_spmm_f32 python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu --others '' > gpurun_out/${TAG}_ncu1.log 2>&1; echo "ncu spmm f32 rc=$?"
      timeout 600 $NCU -k 'regex:^k_spmm$' -s 3 -o gpurun_out/${TAG}_prof_spmm_f64 python bench.py --workload k64f64 --steps 2 --warmup 1 --skip-e2e --skip-cpu --others '' > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu spmm f64 rc=$?"
      timeout 600 $NCU -k 'regex:^k_spmv$' -s 3 -o gpurun_out/${TAG}_prof_spmv python bench.py --workload cfg2 --steps 2 --warmup 1 --skip-e2e --skip-cpu --others '' > gpurun_out/${TAG}_ncu3.log 2>&1; echo "ncu spmv rc=$?"
      timeout 600 $NCU -k 'regex:^k_radix_scatter$' -s 4 -o gpurun_out/${TAG}_prof_radix python bench.py --workload cfg4 --steps 2 --warmup 1 --skip-e2e --skip-cpu --others '' > gpurun_out/${TAG}_ncu4.log 2>&1; echo "ncu radix rc=$?"
      ;;
  esac
done
ls -la gpurun_out | tail -25
