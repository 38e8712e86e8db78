python -m pytest tests/test_omni_attention_gpu.py -x -q -m gpu 2>&1 | tail -3
echo default; QUICK=1 NO_SDPA=1 python tools/bench_attention.py
for v in "$@"; do echo $v; UNIGEN_B200_LIB=build/exp/lib_$v.so QUICK=1 NO_SDPA=1 python tools/bench_attention.py; done
python tools/attn_timeline.py > gpurun_out/timeline.txt
