for b in 8 16; do CTB_KERNEL_US=1 timeout 60 python tools/stress_decode.py $b 512 2 2>&1 | tail -2; done
timeout 150 python -m pytest tests/test_gpu_gpt.py -x -q --timeout 80 -k "oracle or every_decode or larger_than" 2>&1 | tail -3
