timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_backbone_gpu.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/exp_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-pixels --no-best-of-n --no-isolated > gpurun_out/exp_benchpy.json 2>/dev/null
