#!/usr/bin/env bash
# The ncu captures behind profiles/rNN_launches.* and profiles/rNN_ncu_summary.md (B200_PROFILING.md recipe).  One GPU.
#   tools/profile_round.sh r02        -> gpurun_out/r02_launches.csv, r02_vit.ncu-rep, r02_cluster.ncu-rep, r02_detok_launches.csv
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
# (1) launch list of one tokenizer step: 3 warm-up steps x 161 launches skipped, the 161 launches of the 4th step captured
#     (206 per step with SETOK_VIT_LN_FOLD=0: SKIP=618 COUNT=206)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled -k regex:setok -s ${SKIP:-483} -c ${COUNT:-161} --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > $OUT/${TAG}_launches.log 2>&1
# (2) ncu --set full of one ViT layer's kernels (LayerNorms folded into the GEMMs: out_proj, fc1, fc2, qkv, attention, ...)
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"gemm_bf16|attn_fullrow|layernorm" -s 40 -c 6 \
    -f -o $OUT/${TAG}_vit python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > $OUT/${TAG}_vit.log 2>&1
# (3) the fused clustering kernel the way the tokenizer runs it (embedded input)
SETOK_DPC_EMBEDDED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:dpc_fused -s 2 -c 1 -f -o $OUT/${TAG}_cluster \
    python tools/run_cluster.py > $OUT/${TAG}_cluster.log 2>&1
# (4) launch list of the detokenizer (config 3's decoder half)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled -k regex:setok --csv \
    --log-file $OUT/${TAG}_detok_launches.csv python tools/bench_detok.py > $OUT/${TAG}_detok_ncu.log 2>&1
ls -la $OUT/${TAG}_*
