#!/bin/bash
# tools/perf_round.sh TAG - kernel-level evidence for the join pipelines (one GPU): GPU tests, per-launch traces of Q3 / Q5 / SSB Q4.x at
# SF100, and ncu --set full captures of k_probe_chain (summarised by tools/ncu_summary.py).  Run ON the GPU box:
#     gpurun --timeout 1500 -- 'bash tools/perf_round.sh r02d'
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python tools/run_queries.py --sf 100 --queries q3,q5,ssb4.1,ssb4.2,ssb4.3 --steps 5 --trace > $OUT/${TAG}_traces.txt 2>&1
echo "traces rc=$?"; grep -v "^   k_agg\|^   k_ht_clear" $OUT/${TAG}_traces.txt | cut -c1-200 | head -150
if [ "${2:-}" = "ncu" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_probe_chain -c 12 -o $OUT/${TAG}_joins -f \
      python tools/run_queries.py --sf 100 --queries q3,q5,ssb4.1 --steps 1 > $OUT/${TAG}_joins_ncu.log 2>&1
  echo "ncu rc=$?"
  [ -f $OUT/${TAG}_joins.ncu-rep ] && python tools/ncu_summary.py $OUT/${TAG}_joins.ncu-rep > $OUT/${TAG}_probe_chain_ncu_full_summary.json
  ls -la $OUT | tail -5
fi
