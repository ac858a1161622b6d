timeout 300 python -m pytest tests/test_gpu_edge.py tests/test_gpu_ssb.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/r02v_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02v_pytest.log
timeout 200 python tools/run_queries.py --sf 100 --queries q3,q5,ssb4.1,ssb4.2,ssb4.3 --trace 2>&1 | grep -v "^{" > gpurun_out/r02v_trace_n1.txt; grep -E "k_prefilter|^trace" gpurun_out/r02v_trace_n1.txt | head -40
CBGPU_PREFILTER_TMA=1 timeout 200 python tools/run_queries.py --sf 100 --queries q3,q5,ssb4.1 --trace 2>&1 | grep -v "^{" > gpurun_out/r02v_trace_n1_tma.txt; grep -E "k_prefilter|^trace" gpurun_out/r02v_trace_n1_tma.txt | head
timeout 500 python bench.py > gpurun_out/r02v_bench_n1.json 2> gpurun_out/r02v_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02v_bench_n1.err
