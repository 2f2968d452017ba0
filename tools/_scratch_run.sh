run() { timeout 300 python tools/sweep.py "$@" 2>&1 | tail -1 | sed -e 's/ place=.*overlap:/ overlap:/' -e 's/ place=.*//' | cut -c1-400; }
export LDPC_B200_OVERLAP=0
for cfg in "1024 8" "1024 6" "1024 4" "896 8" "896 6" "896 4" "768 8"; do set -- $cfg
  export LDPC_B200_OV_THREADS=$1 LDPC_B200_PRODUCERS=$2; echo "c2 threads=$1 producers=$2"; run 10000 3 6 0.38,0.39,0.40,0.41,0.42,0.43 2000000; done
