python tools/sanitize.py > gpurun_out/san_plain.log 2>&1; tail -2 gpurun_out/san_plain.log
timeout 1500 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/san_memcheck.log 2>&1; tail -4 gpurun_out/san_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck python tools/sanitize.py > gpurun_out/san_racecheck.log 2>&1; tail -4 gpurun_out/san_racecheck.log
timeout 900 compute-sanitizer --tool synccheck python tools/sanitize.py > gpurun_out/san_synccheck.log 2>&1; tail -4 gpurun_out/san_synccheck.log
