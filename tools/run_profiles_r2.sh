#!/bin/bash
# Round-2 evidence in one GPU call: GPU tests, the bench line, the ncu launch list of the bench command and one `ncu --set full`
# capture per dominant kernel.  Outputs under gpurun_out/; tools/ncu_summary.py / ncu_traffic.py / launch_summary.py turn them into profiles/.
set -x
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_pytest_final.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2500 -c 1500 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 3 > gpurun_out/r2_launches_bench.log 2>&1
timeout 300 $NCU -k regex:gemm_tc_kernel --launch-skip 2 -c 1 -o gpurun_out/ncu_gemm_swiglu_r2 python tools/prof_gemm.py gate_up > gpurun_out/ncu1.log 2>&1
timeout 300 $NCU -k regex:gemm_tc_kernel --launch-skip 2 -c 1 -o gpurun_out/ncu_gemm_downres_r2 python tools/prof_gemm.py down_res > gpurun_out/ncu2.log 2>&1
timeout 300 $NCU -k regex:omni_attention --launch-skip 3 -c 1 -o gpurun_out/ncu_attn_r2 python tools/prof_attention.py > gpurun_out/ncu3.log 2>&1
SCALE=0.02 timeout 300 $NCU -k regex:decode_fast --launch-skip 3 -c 1 -o gpurun_out/ncu_decode_r2 python tools/prof_decode.py > gpurun_out/ncu_decode_r2.log 2>&1
SCALE=0.02 timeout 300 $NCU -k regex:select_kernel --launch-skip 3 -c 1 -o gpurun_out/ncu_select_r2 python tools/prof_decode.py > gpurun_out/ncu_select_r2.log 2>&1
NO_SDPA=1 timeout 300 python tools/bench_attention.py > gpurun_out/r2_attention_sweep.log 2>&1; cp gpurun_out/attention_sweep.json gpurun_out/r2_attention_sweep.json
timeout 300 python tools/bench_decode.py > gpurun_out/r2_decode_sweep.log 2>&1
