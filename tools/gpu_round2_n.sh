python tools/gpu_tile_sweep.py C1 131072 "0:0,7:0,3:0,7:-1,5:-1,3:-1,2:-1,1:-1"
python tools/gpu_tile_sweep.py C2 37888 "0:0,12:0,9:0,7:0"
