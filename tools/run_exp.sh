: > gpurun_out/exp_bench.log
for L in ml_unigen_b200/csrc/libunigen_b200.so build/exp/lib_*.so; do
  echo $L >> gpurun_out/exp_bench.log; UNIGEN_B200_LIB=$L QUICK=1 NO_SDPA=1 timeout 300 python tools/bench_attention.py >> gpurun_out/exp_bench.log 2>&1
done
