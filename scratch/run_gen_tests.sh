#!/bin/bash
# one process per kernel family: a sticky CUDA error in one must not hide the others
rm -f gpurun_out/gen_ops.log
for k in fprop first_layer dgrad wgrad rejects; do
  echo "=== $k" >> gpurun_out/gen_ops.log
  timeout 60 python -m pytest tests/test_gpu_vgg_tc_ops.py -q -m gpu -x -k $k 2>&1 | grep -v "^$" | head -70 >> gpurun_out/gen_ops.log
done
timeout 120 python -m pytest tests/test_gpu_vgg_tc_plan.py -q -m gpu -s 2>&1 | grep -v "^$" | head -120 > gpurun_out/gen_model.log
timeout 100 python scratch/vgg_time.py 1024 bf16 > gpurun_out/vgg_bf16_1024.log 2>&1
grep -E "===|passed|failed|Error|assert" gpurun_out/gen_ops.log | head -40
grep -E "passed|failed|Error|assert|step|fp32|bf16" gpurun_out/gen_model.log | head -30
tail -20 gpurun_out/vgg_bf16_1024.log
