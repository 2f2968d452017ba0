timeout 120 python tools/sweep.py 10000 3 6 0.38,0.39,0.40,0.41,0.42,0.43 1000000 2>&1 | cut -c1-200
timeout 120 python tools/sweep.py 12000 3 6 0.40,0.42 1000000 2>&1 | cut -c1-200
LDPC_B200_V2C_RES=2 timeout 120 python tools/sweep.py 12000 3 6 0.40,0.42 1000000 2>&1 | cut -c1-200
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
