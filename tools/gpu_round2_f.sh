set -x
python tools/gpu_debug_v3.py > gpurun_out/debug_v3.txt 2>&1; cat gpurun_out/debug_v3.txt
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/gpu_debug_v3.py 4 > gpurun_out/racecheck.txt 2>&1; grep -v "^warps" gpurun_out/racecheck.txt | head -60
timeout 600 compute-sanitizer --tool initcheck --print-limit 20 python tools/gpu_debug_v3.py 4 > gpurun_out/initcheck.txt 2>&1; grep -v "^warps" gpurun_out/initcheck.txt | head -60
