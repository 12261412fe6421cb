python tools/gpu_tile_sweep.py C2 37888 "0:0,15:4,23:0,23:4,23:8,31:4,31:8,15:-1"
python tools/gpu_tile_sweep.py C5 37888 "0:0,23:0,23:4,15:4"
python tools/gpu_tile_sweep.py C4 18944 "0:0,15:-1,15:16"
python -m pytest tests -m gpu -q --timeout 1200 -x 2>&1 | tail -5
