#!/bin/bash
# Per-kernel SASS evidence of the Blackwell data path: counts of UBLKCP (cp.async.bulk, the TMA unit), UTMALDG (cp.async.bulk.tensor: tensor-map TMA), FFMA2 (packed fp32), SYNCS
# (mbarrier), LDGSTS (cp.async), LDG / STG, and the library-wide absence of getenv.  Runs without a GPU.
#   profiles/tools/sass_summary.sh > profiles/r2/sass_tma.txt
set -e
LIB=${1:-hyperspectral_super-resolution_b200/csrc/libhsr_b200.so}
echo "# $(basename $LIB): $(cuobjdump -lelf $LIB | head -3 | tr '\n' ' ')"
echo "# undefined host symbols that would read the environment: $(nm -D --undefined-only $LIB | grep -c -w getenv || true) (getenv)"
cuobjdump -sass $LIB | awk '
/Function :/ { name=$3; order[++n]=name }
/UBLKCP/ { ublkcp[name]++ }
/UTMALDG/ { utma[name]++ }
/FFMA2/ { ffma2[name]++ }
/SYNCS/ { syncs[name]++ }
/LDGSTS/ { ldgsts[name]++ }
/ LDG\./ { ldg[name]++ }
/ STG\./ { stg[name]++ }
/ LDS/ { lds[name]++ }
/ STS/ { sts[name]++ }
END {
  printf "%-8s %-8s %-8s %-8s %-8s %-6s %-6s %-6s %-6s  %s\n", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "FFMA2", "LDG", "STG", "LDS", "STS", "kernel (mangled)"
  for (i = 1; i <= n; ++i) { k = order[i];
    printf "%-8d %-8d %-8d %-8d %-8d %-6d %-6d %-6d %-6d  %s\n", ublkcp[k], utma[k], syncs[k], ldgsts[k], ffma2[k], ldg[k], stg[k], lds[k], sts[k], k }
}'
