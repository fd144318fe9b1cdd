python tests/prof_step.py 1000000 0.3 10 2>&1 | grep -E "wall|interp|spread|wave_bin"
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
