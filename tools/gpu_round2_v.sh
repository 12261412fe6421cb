for c in 0 100; do echo "carveout $c"; G3_CARVEOUT=$c G3CUDA_LIBRARY=gadget3_b200/csrc/variants/libg3cuda_carve.so python tools/gpu_tile_sweep.py C2 37888 "0:0"; done
python tools/gpu_tile_sweep.py C2 37888 "0:0"
