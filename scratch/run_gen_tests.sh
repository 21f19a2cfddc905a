#!/bin/bash
# one process per kernel family: a sticky CUDA error in one must not hide the others
for k in fprop first_layer dgrad wgrad rejects; do
  echo "=== $k" >> gpurun_out/gen_ops.log
  timeout 60 python -m pytest tests/test_gpu_vgg_tc_ops.py -q -m gpu -x -k $k 2>&1 | grep -v "^$" | head -70 >> gpurun_out/gen_ops.log
done
timeout 90 python -m pytest tests/test_gpu_vgg_tc_plan.py -q -m gpu -x 2>&1 | grep -v "^$" | head -90 > gpurun_out/gen_model.log
grep -E "===|passed|failed|Error|assert" gpurun_out/gen_ops.log | head -40
grep -E "passed|failed|Error|assert" gpurun_out/gen_model.log | head -20
