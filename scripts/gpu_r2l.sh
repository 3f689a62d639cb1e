#!/bin/bash
# round 2, call L (2 GPUs): hygiene -- racecheck of the Cholesky kernels with the full report, initcheck + memcheck of the
# peer-memory exchange on 2 GPUs (quick worker), ncu of rotate_tc / jacobi; then the final 1-GPU validation (pytest, smoke, bench).
TAG=${1:-r2l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
echo "== racecheck, Cholesky kernels (full report)"
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_parity.py -q -m gpu -x \
   -k "tiled_cholesky and synth_implicit_cg_k128 and 0-0" > $OUT/racecheck_chol_detail.txt 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" $OUT/racecheck_chol_detail.txt | tail -3
grep -E "hazard detected|and (Write|Read) access at|Current Value" $OUT/racecheck_chol_detail.txt | sort | uniq -c | sort -rn | head -12
echo "== memcheck + initcheck of the peer-memory exchange (2 GPUs, quick worker)"
for TOOL in memcheck initcheck; do
  B200ALS_WORKER_QUICK=1 B200ALS_EXCHANGE=p2p timeout 900 compute-sanitizer --tool $TOOL --target-processes all \
     python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 tests/multigpu_worker.py > $OUT/sanitizer_p2p_$TOOL.txt 2>&1
  grep -E "ERROR SUMMARY|MULTIGPU_OK|kernel 3 world" $OUT/sanitizer_p2p_$TOOL.txt | tail -5 | cut -c1-200
done
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke.txt
echo "== bench c2 / c3-chol (roofline peaks)"
for WL in c2 c3-chol; do timeout 300 python bench.py --workload $WL --steps 3 2>&1 | tail -1 | tee $OUT/bench_$WL.json | cut -c1-200; done
ls $OUT
