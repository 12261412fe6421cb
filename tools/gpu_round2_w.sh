python tools/gpu_tile_sweep.py C2 37888 "0:0"
python tools/gpu_tile_sweep.py C4 18944 "0:0"
python tools/gpu_tile_sweep.py C5 37888 "0:0"
python tools/gpu_tile_sweep.py C1 65536 "0:0"
python tools/gpu_phase_clocks.py gadget3_b200/csrc/variants/libg3cuda_phase.so C2 4736
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x 2>&1 | tail -4
