#!/bin/bash
# GPU: times every tools/variants/libsnvknn_${PREFIX}*.so on the cfg-2 shape (296 windows) and the cfg-5 shard shape with the engines in
# $ENGINES.  Output: gpurun_out/${TAG}_variants.txt
cd "$(dirname "$0")/.."
TAG=${TAG:-r2}; PREFIX=${PREFIX:-dg_}; ENGINES=${ENGINES:-"tc4x2 tc4x2ta"}
mkdir -p gpurun_out
{
for lib in rag_snvbert_b200/libsnvknn.so tools/variants/libsnvknn_${PREFIX}*.so; do
  [ -f "$lib" ] || continue
  for e in $ENGINES; do
    echo "== $(basename $lib) $e cfg2"; SNVKNN_LIB=$PWD/$lib SNV_HAMMING_ENGINE=$e W=296 timeout 90 python tools/time_hamming.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('kernel_ms %.4f ms %.4f checksum %d'%(d['kernel_ms'],d['ms'],d['checksum']))"
    [ -n "$CFG5" ] && { echo "== $(basename $lib) $e cfg5 shard"; SNVKNN_LIB=$PWD/$lib SNV_HAMMING_ENGINE=$e W=8 N=25000 Q=10000 K=32 timeout 90 python tools/time_hamming.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('kernel_ms %.4f ms %.4f checksum %d'%(d['kernel_ms'],d['ms'],d['checksum']))"; }
  done
done
} > gpurun_out/${TAG}_variants.txt 2>&1
cat gpurun_out/${TAG}_variants.txt
