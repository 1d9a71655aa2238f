python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "tcgen05_path or dgrad_bn_fusion" 2>&1 | tail -2
cd tests
for v in default warpwide; do
  if [ $v = warpwide ]; then export SVK_LIB_PATH=$PWD/../pytorch-kaldi-resnet_b200/svk/libsvk_warpwide.so; else unset SVK_LIB_PATH; fi
  echo "== $v"
  SVK_PROF=1 python prof_conv.py 256 32 2>&1 | grep -E "^(fwd|dgr-bn) +\(40, 200, 32, 32" | cut -c1-130
  SVK_PROF=1 python prof_conv.py 256 64 2>&1 | grep -E "^(fwd|dgr-bn) +\(20, 100, 64, 64" | cut -c1-130
done
