for v in faithful rcplib both; do echo "== $v"; G3CUDA_LIBRARY=gadget3_b200/csrc/variants/libg3cuda_$v.so python tools/gpu_debug_v3.py 2>&1 | head -3; done
