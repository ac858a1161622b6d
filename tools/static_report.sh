#!/bin/bash
# tools/static_report.sh > profiles/<tag>_static_kernels.txt - what can be said about the built kernels without a GPU:
# registers / stack / shared memory per kernel (cuobjdump -res-usage) and, for the two hot kernels, the SASS mnemonics that
# show how they move data (UBLKCP = cp.async.bulk, SYNCS = mbarrier, LDG.E.128 = 16-byte loads, ATOMS / RED = atomics).
set -u
cd "$(dirname "$0")/.."
SO=cloudberry_b200/libcbgpu.so
echo "# $(nvcc --version | tail -2 | head -1); sm_100a; $(date -u +%F)"
echo "## resources per kernel (cuobjdump -res-usage $SO)"
cuobjdump -res-usage $SO 2>/dev/null | awk '/Function/ {f=$2} /REG:/ {print f, $0}' | sed 's/:  */ /' | c++filt | sort
for pat in 'k_scan_agg_smallILi4ELi55ELb1ELi1' 'k_probe_chain'; do
    echo
    echo "## SASS mnemonic histogram: $pat"
    cuobjdump -sass $SO 2>/dev/null | awk -v pat="$pat" '
        /Function :/ {on = index($0, pat) > 0}
        on && /^ +\/\*[0-9a-f]+\*\/ / {m = $2; if (substr(m, 1, 1) == "@") m = $3; sub(/;$/, "", m); n[m]++; tot++}
        END {print tot, "instructions"; for (k in n) if (k ~ /^(UBLKCP|SYNCS|LDG|LD\.|LDS|LDSM|STG|ST\.|STS|ATOM|RED|CCTL|MEMBAR|FENCE|BAR|WARPSYNC|SHFL|VOTE|MATCH|REDUX|IMAD\.WIDE|DMUL|DFMA|MUFU)/) print n[k], k}' | sort -rn
done
