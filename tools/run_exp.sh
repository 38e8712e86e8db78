timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/exp_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-pixels --no-best-of-n > gpurun_out/exp_benchpy.json 2>/dev/null
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/exp_smoke.log 2>&1
