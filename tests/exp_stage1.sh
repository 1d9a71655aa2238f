cd tests
for g in 2 3; do for s in 0 1; do
  echo "== groups=$g single_fwd32=$s"
  SVK_PROF=1 SVK_GATHER3_GROUPS=$g SVK_SINGLE_HALO_FWD32=$s python prof_conv.py 256 32 2>&1 | grep -E "^fwd |^fwd-ns" | cut -c1-200
done; done
