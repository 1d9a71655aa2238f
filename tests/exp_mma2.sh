# per-role counters of the resident-filter conv kernels: one or two MMA-issuing warps x two or three epilogue groups
C=${1:-32}
for m in 1 2; do for g in 2 3; do
  echo "== MMA warps $m, epilogue groups $g"
  SVK_GATHER3_MMA=$m SVK_GATHER3_GROUPS=$g SVK_PROF=1 python tests/prof_conv.py 256 $C 2>&1 | grep -v "^op" | grep "3, 1)" | cut -c1-150
done; done
