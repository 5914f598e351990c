#!/usr/bin/env bash
# compute-sanitizer over the hand-rolled mbarrier / TMEM / TMA pipelines: memcheck, racecheck and synccheck
# on a small set of GPU tests that drive every kernel family once (the sanitizer slows kernels 10-100x, so the full-size
# configurations are left to the plain test run).  Usage (on the GPU box):  tools/sanitize.sh [out_dir]
set -u
OUT=${1:-gpurun_out/sanitize}
mkdir -p "$OUT"
CS=${COMPUTE_SANITIZER:-/usr/local/cuda/bin/compute-sanitizer}
TESTS="tests/test_gpu_kernels.py::test_attention_vit_hd64[17-3-2] tests/test_gpu_kernels.py::test_attention_vit_hd64[82-2-2] \
tests/test_gpu_kernels.py::test_attention_vit_hd64[257-3-16] tests/test_gpu_kernels.py::test_layernorm \
tests/test_gpu_dpc.py::test_dpc_golden_bit_exact tests/test_gpu_model.py::test_head_golden tests/test_gpu_model.py::test_tower_golden_small \
tests/test_gpu_preprocess.py::test_golden_cases_bit_exact tests/test_gpu_splice.py tests/test_gpu_detok.py::test_detok_golden \
tests/test_gpu_kernels.py::test_gemm_epilogues"
# both GEMM epilogues (row-owner / transposing) on the small shapes
KEXPR="(row_owner and (391 or 77-96 or 1-32)) or (fold_chain and (65-128 or 160-96 or 300-1024))"
# the LayerNorm-fold tower (both producing epilogues, the fused pre-LN record) and the vector im2col
MEXPR="fold_matches or im2col_vector"
rc=0
for tool in memcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  # shellcheck disable=SC2086
  timeout 1500 "$CS" --tool "$tool" --error-exitcode 9 --launch-timeout 120 \
      python -m pytest $TESTS -m gpu -q -x -p no:cacheprovider > "$OUT/$tool.log" 2>&1
  r=$?
  timeout 900 "$CS" --tool "$tool" --error-exitcode 9 --launch-timeout 120 \
      python -m pytest tests/test_gpu_kernels.py -k "$KEXPR" -m gpu -q -x -p no:cacheprovider >> "$OUT/$tool.log" 2>&1
  r2=$?
  [ $r2 -ne 0 ] && r=$r2
  timeout 900 "$CS" --tool "$tool" --error-exitcode 9 --launch-timeout 120 \
      python -m pytest tests/test_gpu_model.py -k "$MEXPR" -m gpu -q -x -p no:cacheprovider >> "$OUT/$tool.log" 2>&1
  r3=$?
  [ $r3 -ne 0 ] && r=$r3
  tail -4 "$OUT/$tool.log"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$OUT/$tool.log" | tail -2
  [ $r -ne 0 ] && rc=$r
done
exit $rc
