python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv_tcgen05 or dgrad_bn or repeatable or epilogues" 2>&1 | tail -2
for C in 32 64; do
for L in head new; do
  if [ $L = head ]; then LP=$PWD/pytorch-kaldi-resnet_b200/svk/libsvk_head.so; else LP=$PWD/pytorch-kaldi-resnet_b200/svk/libsvk.so; fi
  echo "== $L $C"; SVK_LIB_PATH=$LP SVK_PROF=1 python tests/prof_conv.py 256 $C 2>&1 | grep "3, 1)" | cut -c1-150
done; done
bash tests/ab_lib.sh head 2
