# usage: bash tests/exp_dbg.sh <channels> <variant>... — per-role cycle counters of the conv kernels for diagnostic builds
C=$1; shift
for v in base "$@"; do
  if [ $v = base ]; then L=$PWD/pytorch-kaldi-resnet_b200/svk/libsvk.so; else L=$PWD/pytorch-kaldi-resnet_b200/svk/libsvk_$v.so; fi
  echo "== $v"
  SVK_LIB_PATH=$L SVK_PROF=1 python tests/prof_conv.py 256 $C 2>&1 | grep -v "^op" | cut -c1-150
done
