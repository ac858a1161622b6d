#!/bin/bash
# tools/ab_round.sh TAG [ncu] - A/B timings of the join pipelines' kernels under the tuning knobs (one GPU), plus an optional ncu capture
set -u
TAG=${1:-rXX}
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
run() { echo "== $1"; shift; env "$@" timeout 200 python tools/run_queries.py --sf 100 --queries q3,q5,ssb4.1,ssb4.2,ssb4.3 --steps 5 --trace 2>&1 | grep "k_probe_chain  *[0-9]*\.[0-9]* ms\|k_ht_build  *[0-9]\.[0-9]* ms\|\"query\"" | awk '{ if ($1 ~ /k_/) { if ($2+0 > 0.3) print "   " $1, $2 } else print substr($0,1,70) }'; }
run default X=1
run no-smem-ht CBGPU_NO_SMEM_HT=1
run no-fuse0 CBGPU_NO_FUSE0=1
if [ "${2:-}" = "ncu" ]; then
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_probe_chain --launch-skip 2 --launch-count 1 -o $OUT/${TAG}_ssb -f \
      python tools/run_queries.py --sf 100 --queries ssb4.1 --steps 1 > $OUT/${TAG}_ssb_ncu.log 2>&1
  echo "ncu rc=$?"
  [ -f $OUT/${TAG}_ssb.ncu-rep ] && python tools/ncu_summary.py $OUT/${TAG}_ssb.ncu-rep > $OUT/${TAG}_ssb_probe_chain_ncu_full_summary.json
  ncu -i $OUT/${TAG}_ssb.ncu-rep --page source --csv > $OUT/${TAG}_ssb_sass.csv 2>/dev/null
  python tools/ncu_lines.py $OUT/${TAG}_ssb_sass.csv cloudberry_b200/libcbgpu.so _Z13k_probe_chain8PcParams 40 > $OUT/${TAG}_ssb_lines.txt 2>&1
fi
