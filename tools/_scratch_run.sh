python tools/anomaly.py 4000 20 1000000 1
python tools/anomaly.py 4000 20 1000000 0
python tools/anomaly.py 1000 1 100000000 1
