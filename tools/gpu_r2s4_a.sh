# session 4, call a: gradient kernel at 15 / 11 / 7 unit warps (512 / 384 / 256-thread instantiations), NTAN 4 / 8 / 2
python tools/gpu_grad_warps.py C2 8192 15,11,7 2>&1 | tail -5
G3CUDA_LIBRARY=$PWD/gadget3_b200/csrc/variants/libg3cuda_ntan8.so python tools/gpu_grad_warps.py C2 8192 15,7 2>&1 | tail -3
G3CUDA_LIBRARY=$PWD/gadget3_b200/csrc/variants/libg3cuda_ntan2.so python tools/gpu_grad_warps.py C2 8192 15,7 2>&1 | tail -3
python tools/gpu_grad_warps.py C1 16384 15,11,7 2>&1 | tail -4
python tools/gpu_grad_warps.py C5 2048 15,11,7 2>&1 | tail -4
