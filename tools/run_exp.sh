ncu --set full --clock-control none --import-source on -f -k regex:gemm_tc_kernel --launch-skip 2 -c 1 -o gpurun_out/ncu_gemm_swiglu_r2 python tools/prof_gemm.py gate_up > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -f -k regex:gemm_tc_kernel --launch-skip 2 -c 1 -o gpurun_out/ncu_gemm_downres_r2 python tools/prof_gemm.py down_res > gpurun_out/ncu2.log 2>&1
