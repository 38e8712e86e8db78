timeout 900 python -m pytest tests/test_decode_step_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/exp_pytest.log
BS=1,16,64,128 timeout 300 python tools/bench_decode.py > gpurun_out/exp_decode.log 2>&1
