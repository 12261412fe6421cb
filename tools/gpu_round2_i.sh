python tools/gpu_tile_sweep.py C2 37888 "0:0,15:4,23:0,23:4,23:8,31:4,31:8,15:-1,12:-1"
python tools/gpu_tile_sweep.py C1 65536 "0:0,23:0,23:4,31:4"
python tools/gpu_tile_sweep.py C5 37888 "0:0,23:0,23:4,31:4"
