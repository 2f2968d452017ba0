nproc
python tools/sweep.py 10000 3 6 0.38,0.39,0.40,0.41,0.42,0.43 1000000 2>&1 | cut -c1-330
python tools/sweep.py 1000 3 6 0.40 20000000 2>&1 | cut -c1-330
python tools/sweep.py 4000 3 6 0.40,0.42 4000000 2>&1 | cut -c1-330
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
