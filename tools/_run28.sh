python tools/sweep.py 1000 3 6 0.40 100000000 2>&1 | cut -c60-175
python tools/sweep.py 10000 3 6 0.38,0.39,0.40,0.41,0.42,0.43 2000000 2>&1 | cut -c60-175
timeout 600 python -m pytest tests -m gpu -x -q -k "staged or golden or fuzz" 2>&1 | tail -2
