#!/bin/bash
# First GPU pass: parity tests in stages (each under its own timeout so a hung kernel cannot
# hold the box), smoke, short bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import torch;print(torch.__version__, torch.cuda.get_device_name(0))" >> gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 25 gpurun_out/$name.log; }
run t1_simple 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=20 -k "minv or mc_ or interp2x or grid_sampler"
run t2_sdf 240 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=20 -k "sdf_small or sdf_full"
run t3_fields 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=20 -k "deformer or render or cardinal or trace or seg3d"
run smoke 240 python __graft_entry__.py --smoke
run bench 600 python bench.py --steps 3 --warmup 3
grep -h '"metric"' gpurun_out/bench.log > gpurun_out/bench_line.json
