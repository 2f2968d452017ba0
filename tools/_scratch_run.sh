timeout 600 python tools/ens_bench2.py 2>&1 | head -8
ncu --set full --clock-control none --import-source on -k regex:e1_mc_kernel --launch-skip 1 -c 1 -f -o gpurun_out/prof_ens python tools/ens_bench2.py > gpurun_out/ncu_ens.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_ens.ncu-rep 12 > gpurun_out/r2i_ens_ncu_full.txt 2>&1
python tools/ncu_regions.py gpurun_out/prof_ens.ncu-rep >> gpurun_out/r2i_ens_ncu_full.txt 2>&1
rm -f gpurun_out/prof_ens.ncu-rep
grep -A22 "regions by" gpurun_out/r2i_ens_ncu_full.txt
