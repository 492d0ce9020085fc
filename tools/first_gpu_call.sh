#!/usr/bin/env bash
# First GPU call of the next session: everything that was written after round 1's GPU minutes were spent, in order of
# increasing risk, each step under its own timeout, logs into gpurun_out/ (merged back by gpurun).
#
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/first_gpu_call.sh'
#
# Reads: DESIGN.md section 10 (status table).  A step that fails does not stop the later ones.
set -u
mkdir -p gpurun_out
log() { echo "=== $*" | tee -a gpurun_out/first_call_summary.log; }
run() {  # run <seconds> <logfile> <command...>
    local t=$1 f=$2; shift 2
    log "$* (timeout ${t}s) -> $f"
    timeout "$t" "$@" > "gpurun_out/$f" 2>&1
    local rc=$?
    log "exit $rc: $(tail -n 1 "gpurun_out/$f" | cut -c1-200)"
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.csv 2>&1
run 300 smoke.log python -c "import __graft_entry__ as g; g.smoke()"
run 600 verified_suite.log python -m pytest tests/test_gpu_parity.py -q
run 300 ref_kernels.log python -m pytest tests/test_zy_reference_cuda_kernels_gpu.py -q -s
run 300 zz_llama.log python -m pytest tests/test_zz_llama_model_gpu.py -q -s
run 300 zz_deepseek.log python -m pytest tests/test_zz_deepseek_block_gpu.py -q -s
run 300 zzv_shapes.log python -m pytest tests/test_zzv_full_size_transform_shapes_gpu.py -q -s
FLATQ_RUN_UNVERIFIED=1 run 300 zzz_grouped.log python -m pytest tests/test_zzz_grouped_gemm_gpu.py -q -s
FLATQ_RUN_UNVERIFIED=1 run 300 zzz_glue.log python -m pytest tests/test_zzz_glue_kernels_gpu.py -q -s
FLATQ_RUN_UNVERIFIED=1 run 300 zzz_decode_split.log python -m pytest tests/test_zzz_decode_split_gpu.py -q -s
run 400 bench_default.json python bench.py --steps 10 --warmup 3
run 120 ref_kernels_bench.json python tests/tools/ref_kernels_bench.py
FLATQ_RUN_UNVERIFIED=1 run 300 decode_attn_bench.jsonl python tests/tools/decode_attn_bench.py
run 600 bench_model.json python bench.py --workload llama8b-model --steps 3 --warmup 3
FLATQ_DECODE_SPLIT=1 run 600 bench_model_split_decode.json python bench.py --workload llama8b-model --steps 3 --warmup 3
FLATQ_DECODE_SPLIT=1 FLATQ_FUSED_GLUE=1 run 600 bench_model_split_decode_fused_glue.json python bench.py --workload llama8b-model --steps 3 --warmup 3
run 400 bench_deepseek_loop.json python bench.py --workload deepseek-block --steps 3 --warmup 3
FLATQ_GROUPED_MOE=1 run 400 bench_deepseek_grouped.json python bench.py --workload deepseek-block --steps 3 --warmup 3
run 300 ncu_launches_deepseek.log ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/ncu_launches_deepseek.csv python bench.py --workload deepseek-block --steps 1 --warmup 1
log "done"
