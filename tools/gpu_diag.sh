#!/bin/bash
# Run every GPU test group in its own process (a device trap must not take the others down) and collect logs.
# usage: tools/gpu_diag.sh [pytest -k expressions...]   (default: a fixed list of groups)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/diag.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv >> $LOG 2>&1
TGROUPS=("$@")
if [ ${#TGROUPS[@]} -eq 0 ]; then
  TGROUPS=("test_gemm_plain" "test_gemm_bias_act_residual or test_gemm_inplace or test_gemm_row_remap or test_gemm_rejects" "test_gemm_silu_gate" "test_layernorm or test_rmsnorm" "test_attention" "test_vq" "test_patchify or test_embedding or test_rope")
fi
for g in "${TGROUPS[@]}"; do
  echo "=== GROUP: $g" >> $LOG
  timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "$g" -p no:cacheprovider 2>&1 | tail -80 >> $LOG
  echo "=== exit: $?" >> $LOG
done
grep -E "^=== GROUP|passed|failed|error" $LOG
