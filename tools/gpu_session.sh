#!/bin/bash
# One gpurun call of the development loop: smoke, GPU parity tests, bench, sweep, ncu launch list + full capture.
# Usage: tools/gpu_session.sh <tag> [steps...]; everything lands in gpurun_out/<tag>_*.
set -u
TAG=${1:-s}
shift || true
WHAT=${*:-"smoke tests bench sweep ncu"}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
for w in $WHAT; do
  case $w in
    smoke) timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" ;;
    tests) timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/${TAG}_tests.log ;;
    bench) timeout 900 python bench.py --steps 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/${TAG}_bench.json ;;
    benchref) timeout 600 python bench.py --impl reference --steps 3 > gpurun_out/${TAG}_benchref.json 2> gpurun_out/${TAG}_benchref.err; echo "benchref rc=$?"; cat gpurun_out/${TAG}_benchref.json ;;
    sweep) timeout 900 python tools/sweep.py > gpurun_out/${TAG}_sweep.jsonl 2> gpurun_out/${TAG}_sweep.err; echo "sweep rc=$?"; tail -3 gpurun_out/${TAG}_sweep.err ;;
    ncufull)
      timeout 600 ncu --set full --clock-control none --import-source on -k k_spmm -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_spmm \
        python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu --others '' > gpurun_out/${TAG}_ncu_spmm.log 2>&1; echo "ncu-spmm rc=$?"
      timeout 600 ncu --set full --clock-control none --import-source on -k k_spmv -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_spmv \
        python bench.py --workload cfg2 --steps 2 --warmup 1 --skip-e2e --skip-cpu --others '' > gpurun_out/${TAG}_ncu_spmv.log 2>&1; echo "ncu-spmv rc=$?"
      timeout 600 ncu --set full --clock-control none --import-source on -k k_spmm -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_spmm_f64 \
        python bench.py --workload k64f64 --steps 2 --warmup 1 --skip-e2e --skip-cpu --others '' > gpurun_out/${TAG}_ncu_spmm64.log 2>&1; echo "ncu-spmm64 rc=$?"
      ;;
    ncu)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_spmm|k_spmv|k_radix|k_csc|k_col|k_expand|k_scan|k_transpose' -c 300 --csv \
        --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu-launches rc=$?"
      timeout 600 ncu --set full --clock-control none --import-source on -k k_spmm -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_spmm \
        python bench.py --steps 2 --warmup 1 --skip-e2e --skip-cpu --others '' > gpurun_out/${TAG}_ncu_spmm.log 2>&1; echo "ncu-spmm rc=$?"
      timeout 600 ncu --set full --clock-control none --import-source on -k k_spmv -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_spmv \
        python bench.py --workload cfg2 --steps 2 --warmup 1 --skip-e2e --skip-cpu --others '' > gpurun_out/${TAG}_ncu_spmv.log 2>&1; echo "ncu-spmv rc=$?"
      ;;
  esac
done
ls -la gpurun_out | tail -20
