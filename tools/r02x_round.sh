timeout 200 python -m pytest tests/test_gpu_edge.py -m gpu -q -x -k prefilter > gpurun_out/r02x_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02x_pytest.log
timeout 200 python tools/run_queries.py --sf 100 --queries q3,q5,ssb4.1 --trace 2>&1 | grep -v "^{" > gpurun_out/r02x_trace_n1.txt; grep -E "k_prefilter|^trace" gpurun_out/r02x_trace_n1.txt | head -40
timeout 500 python bench.py > gpurun_out/r02x_bench_n1.json 2> gpurun_out/r02x_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02x_bench_n1.err
