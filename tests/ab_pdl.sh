for i in 1 2; do
for v in nopdl A B C; do
  case $v in
    nopdl) export SVK_DISABLE_PDL=1; unset SVK_LIB_PATH;;
    A) unset SVK_DISABLE_PDL; export SVK_LIB_PATH=$PWD/pytorch-kaldi-resnet_b200/svk/libsvk_pdlA.so;;
    B) unset SVK_DISABLE_PDL; unset SVK_LIB_PATH;;
    C) unset SVK_DISABLE_PDL; export SVK_LIB_PATH=$PWD/pytorch-kaldi-resnet_b200/svk/libsvk_pdlC.so;;
  esac
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
done; done
