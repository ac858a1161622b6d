#!/bin/bash
# tools/gpu_round.sh TAG - one gpurun call's worth of evidence for profiles/ (run it ON the GPU box, one GPU):
#     gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r02a'
# then, back in the container:  cp gpurun_out/TAG_* profiles/   (gpurun_out/ is scratch, profiles/ is tracked)
#
#   1. pytest -m gpu                                   -> gpurun_out/TAG_pytest.log
#   2. bench.py, default flags (never under a profiler) -> gpurun_out/TAG_bench_n1.json
#   3. ncu launch list of a short bench.py run          -> gpurun_out/TAG_bench_launches.csv
#   4. ncu --set full of the Q1 scan+agg kernel (and from it TAG_q1_kernel_traffic.json = bench.py's roofline.traffic record) and of the join pipelines' probe chain, summarised by tools/ncu_summary.py
#                                                       -> gpurun_out/TAG_q1_scan_agg_ncu_full_summary.json,
#                                                          gpurun_out/TAG_q3q5_probe_chain_ncu_full_summary.json
# Every step has its own timeout so that a hang costs minutes, not the box.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
cd "$(dirname "$0")/.."

timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log

timeout 600 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
echo "bench rc=$?"
cut -c1-600 $OUT/${TAG}_bench_n1.json

timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/${TAG}_bench_under_ncu.log 2>&1
echo "launch list rc=$?"

timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_agg_small -c 2 -o $OUT/${TAG}_q1 -f \
    python tools/run_queries.py --sf 100 --queries q1 --steps 1 > $OUT/${TAG}_q1_ncu.log 2>&1
echo "ncu q1 rc=$?"
[ -f $OUT/${TAG}_q1.ncu-rep ] && python tools/ncu_summary.py $OUT/${TAG}_q1.ncu-rep > $OUT/${TAG}_q1_scan_agg_ncu_full_summary.json
# the bench line's roofline.traffic: measured on THIS source (sha256 inside); copy to profiles/q1_kernel_traffic.json
[ -s $OUT/${TAG}_q1_scan_agg_ncu_full_summary.json ] && python tools/q1_traffic_json.py $OUT/${TAG}_q1_scan_agg_ncu_full_summary.json > $OUT/${TAG}_q1_kernel_traffic.json

timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_probe_chain -c 8 -o $OUT/${TAG}_joins -f \
    python tools/run_queries.py --sf 100 --queries q3,q5 --steps 1 > $OUT/${TAG}_joins_ncu.log 2>&1
echo "ncu joins rc=$?"
[ -f $OUT/${TAG}_joins.ncu-rep ] && python tools/ncu_summary.py $OUT/${TAG}_joins.ncu-rep > $OUT/${TAG}_q3q5_probe_chain_ncu_full_summary.json
ls -la $OUT | tail -20
