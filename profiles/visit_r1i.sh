timeout 600 python -m pytest tests/test_gpu_tc_gemm.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python profiles/bench_softmax_grad.py 2>&1 | tee gpurun_out/r1i_softmax_grad.log
