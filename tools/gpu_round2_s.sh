G3CUDA_LIBRARY=gadget3_b200/csrc/variants/libg3cuda_kfirst.so python tools/gpu_tile_sweep.py C2 37888 "0:0"
python tools/gpu_tile_sweep.py C2 37888 "0:0"
python tools/gpu_phase_clocks.py gadget3_b200/csrc/variants/libg3cuda_kfirst_phase.so C2 4736
