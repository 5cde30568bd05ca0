#!/bin/bash
# Run the GPU parity suites in separate processes (a trapped kernel kills its CUDA context) with hard timeouts.
# Usage (on the GPU box): bash scripts/gpu_checks.sh [suite ...]      logs -> gpurun_out/
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
run() {  # name, timeout, args...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$t" python -m pytest -q -m gpu -p no:cacheprovider --timeout 240 "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit=$? $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
suites=${@:-"probe gemm bandwidth attn_fwd"}
for s in $suites; do
  case $s in
    probe)
      for k in "64-0" "128-0" "64-1" "128-1"; do run "probe_$k" 200 tests/test_gpu_primitives.py -k "probe_ts and $k"; done
      # test ids are [K-mode]
      ;;
    gemm)
      run gemm_nt 300 tests/test_gpu_primitives.py -k "gemm_layouts and False-False"
      run gemm_nn 300 tests/test_gpu_primitives.py -k "gemm_layouts and False-True"
      run gemm_tn 300 tests/test_gpu_primitives.py -k "gemm_layouts and True-False"
      run gemm_tt 300 tests/test_gpu_primitives.py -k "gemm_layouts and True-True"
      run gemm_epi 300 tests/test_gpu_primitives.py -k "gemm_epilogues"
      ;;
    bandwidth) run bandwidth 600 tests/test_gpu_bandwidth.py ;;
    attn_fwd)
      run attn_fwd_small 300 tests/test_gpu_attention.py -k "forward and (B1T128H1 or B1T256H2)"
      run attn_fwd 600 tests/test_gpu_attention.py -k "bridge_attention_forward or golden or scatter"
      run attn_vit 300 tests/test_gpu_attention.py -k "vit"
      ;;
    attn_bwd) run attn_bwd 600 tests/test_gpu_attention.py -k "backward" ;;
    attn_prop) run attn_prop 600 tests/test_gpu_attention.py -k "properties or property" ;;
    model) run model 900 tests/test_gpu_model.py ;;
    vision) run vision 600 tests/test_gpu_vision.py ;;
    all) run all 1800 tests ;;
  esac
done
cat gpurun_out/summary.txt
