sed -i 's/for name, B in (("C1", 8192), ("C2", 4096), ("C4", 512), ("C5", 2048)):/for name, B in (("C2", 4096),):/' tools/gpu_grad_sweep.py
for v in ntan2 ntan8; do echo "== $v"; G3CUDA_LIBRARY=gadget3_b200/csrc/variants/libg3cuda_$v.so python tools/gpu_grad_sweep.py 2>&1 | grep "kernel 4\|agree"; done
echo "== ntan4"; python tools/gpu_grad_sweep.py 2>&1 | grep "kernel 4\|agree"
