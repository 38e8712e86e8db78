timeout 600 python -m pytest tests/test_omni_attention_gpu.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/exp_pytest.log
NO_SDPA=1 timeout 300 python tools/bench_attention.py > gpurun_out/exp_bench.log 2>&1
