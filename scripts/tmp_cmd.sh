for s in 1 2; do echo "--- slack $s"; MSDA_B200_PACE_SLACK=$s python scripts/time_batch_scaling.py 2>&1 | tail -6; done
python -m pytest tests/test_cuda_parity.py -x -q -m gpu 2>&1 | tail -2
