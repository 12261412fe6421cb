python tools/gpu_tile_sweep.py C2 37888 "0:0"
python tools/gpu_tile_sweep.py C4 18944 "0:0"
python tools/gpu_tile_sweep.py C5 37888 "0:0"
python tools/gpu_tile_sweep.py C1 65536 "0:0"
timeout 600 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6
python tools/gpu_parity_diag.py 2>&1 | grep "kernel 4" | cut -c1-150
