# one GPU round: every stage under its own timeout, all output under gpurun_out/
mkdir -p gpurun_out
timeout 240 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke.log 2>&1
timeout 700 python -X faulthandler -m pytest tests -m gpu -x -q -o faulthandler_timeout=150 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1
timeout 200 python tools/prof_train_kernels.py > gpurun_out/train_kernels.log 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/train_step_once.py > gpurun_out/train_once.log 2>&1
grep -v "Warn\|WeightNorm" gpurun_out/smoke.log | tail -4; tail -4 gpurun_out/pytest.log; tail -1 gpurun_out/bench.log | cut -c1-3000; tail -3 gpurun_out/train_kernels.log
