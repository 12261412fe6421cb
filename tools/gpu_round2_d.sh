set -x
python tools/gpu_parity_diag.py > gpurun_out/parity_diag.txt 2>&1
for k in 3 4; do python bench.py --steps 2 --warmup 1 --no-cpu-baseline --config C3 --kernel $k > gpurun_out/bench_v2_C3_k$k.json 2> gpurun_out/bench_v2_C3_k$k.err; done
python -m pytest tests -m gpu -q --timeout 1200 -x --deselect tests/test_gpu_parity.py::test_nll_and_components --deselect tests/test_gpu_parity.py::test_ragged_batch_larger_than_one_wave > gpurun_out/pytest_rest.txt 2>&1
tail -5 gpurun_out/pytest_rest.txt
cat gpurun_out/parity_diag.txt
