#!/bin/bash
# GPU box: correctness of decoder build variants (conv op tests + the model tests through MOGE_B200_LIB), then a same-box A/B.
# usage: bash tools/ab_decoder.sh variant [variant ...]      (libraries built with MG_VARIANT=<name> moge_b200/csrc/build.sh)
mkdir -p gpurun_out
last=""
for v in "$@"; do
  MOGE_B200_LIB=$PWD/moge_b200/_lib/libmoge_b200_$v.so timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -k conv -x > gpurun_out/abdec_ops_$v.log 2>&1
  echo "== conv op tests [$v] exit $?"; tail -n 2 gpurun_out/abdec_ops_$v.log
  last=$v
done
MOGE_B200_LIB=$PWD/moge_b200/_lib/libmoge_b200_$last.so timeout 600 python -m pytest tests/test_gpu_model.py -q -m gpu -x > gpurun_out/abdec_model_$last.log 2>&1
echo "== model tests [$last] exit $?"; tail -n 3 gpurun_out/abdec_model_$last.log
timeout 900 python tools/ab_variants.py base "$@" --rounds 2 --names conv3x3.res_a.neck.l3,conv3x3.res_b.neck.l3,conv3x3.res_a.neck.l2,conv3x3.res_b.neck.l2,conv3x3up2,conv3x3.post > gpurun_out/abdec_ab.log 2>&1
echo "== ab exit $?"; grep -v "^{" gpurun_out/abdec_ab.log
