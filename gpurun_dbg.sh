CTB_MEGA_TRACE=1 timeout 60 python tools/stress_decode.py 1 512 2 2>&1 | tail -3
timeout 100 python -m pytest tests/test_gpu_gpt.py -x -q --timeout 60 -k "fixture or oracle or every" 2>&1 | tail -2
