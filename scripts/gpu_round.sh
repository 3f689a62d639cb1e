#!/bin/bash
# One gpurun call = tests + bench + profiles (box acquisition dominates the charge, so batch).
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag>'
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/nvsmi.txt 2>&1
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench c3 default"; timeout 900 python bench.py 2>&1 | tail -3 | tee $OUT/bench_c3.json | cut -c1-2600
for WL in c4 c2; do
  echo "== bench $WL (side workload)"; timeout 600 python bench.py --workload $WL --steps 5 2>&1 | tail -1 | tee $OUT/bench_$WL.json | cut -c1-900
done
if [ "$FULL" = "1" ]; then for K in 2 1; do
  echo "== bench c3 kernel=$K"; timeout 600 python bench.py --kernel $K --steps 3 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_c3_kernel$K.json | cut -c1-900
done; fi
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json | cut -c1-1200
echo "== ncu launch list (same command as the bench, fewer steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/launches_bench.log 2>&1
echo "== ncu dram bytes of the CG kernel at C3 size"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:als_cg_resident -s 3 -c 1 --csv --log-file $OUT/resident_dram_c3.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/resident_dram_c3.log 2>&1
echo "== ncu --set full of the CG kernel (1M-user slice)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:als_cg_resident -s 3 -c 1 -f -o $OUT/prof_resident \
    python bench.py --workload c3-small --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/prof_resident.log 2>&1
echo "== ncu --set full of the tiled Cholesky kernel (C2)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:als_chol_tile -s 1 -c 1 -f -o $OUT/prof_chol \
    python bench.py --workload c2 --steps 1 --warmup 3 > $OUT/prof_chol.log 2>&1
echo "== ncu --set full of rotate / gram / jacobi (1M-user slice)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rotate_tc|gram_tc|jacobi" -s 6 -c 4 -f -o $OUT/prof_aux \
    python bench.py --workload c3-small --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/prof_aux.log 2>&1
ls -la $OUT
