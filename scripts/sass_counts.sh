#!/bin/bash
# Per-kernel counts of the SASS instructions that prove what the hot kernels execute (no GPU needed):
#   DMMA    = FP64 tensor-core MMA (mma.sync.m8n8k4.f64)
#   UBLKCP  = bulk asynchronous copy (cp.async.bulk, the TMA engine); SYNCS = mbarrier operations
#   LDGSTS  = cp.async (Ampere-style staged copy);  ATOMG.E.ADD.F64 with destination RZ = fire-and-forget FP64 reduction (red.global.add.f64)
# usage: scripts/sass_counts.sh > profiles/r2_sass_counts.txt   (after the library is built)
cd "$(dirname "$0")/../magmaclustr_b200/lib" || exit 1
for o in tasks3 estep3 mstep3 pairs3 dense3 dense_big bigtasks covbuild; do
  echo "== $o.o"
  cuobjdump -sass $o.o 2>/dev/null | awk '
    /Function :/ {fn=$3}
    /DMMA/ {d[fn]++; seen[fn]=1}
    /UBLKCP/ {u[fn]++; seen[fn]=1}
    /SYNCS/ {s[fn]++; seen[fn]=1}
    /LDGSTS/ {l[fn]++; seen[fn]=1}
    /ATOMG.*ADD\.F64.*RZ, desc|RED.*F64/ {r[fn]++; seen[fn]=1}
    END {for (f in seen) printf "%-72s DMMA %5d  UBLKCP %3d  SYNCS(mbarrier) %3d  LDGSTS %3d  RED.F64 %3d\n", substr(f,1,72), d[f], u[f], s[f], l[f], r[f]}' | sort
done
