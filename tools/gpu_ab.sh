#!/bin/bash
# A/B timing of library variants (variants/libiactrace_b200_*.so built by tools/build_variant.sh) against the product build.
# usage: tools/gpu_ab.sh "workload ..." [steps]
mkdir -p gpurun_out
WL=${1:-"ct5_point_4096x115_hex ct5_point_4096x115_square ct3_matrix_64x64_M64 ct3_matrix_64x64_M1000 cassegrain_1e9"}
STEPS=${2:-30}
run() { IACTRACE_B200_LIB=$2 python bench.py --workload $1 --steps $STEPS --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-28s %-12s %8.3f ms %7.2f Grays/s' % ('$1', '$3', d['ms_per_step'], d['value']/1e9))"; }
for w in $WL; do
  run $w "" product
  for v in variants/libiactrace_b200_*.so; do
    [ -f "$v" ] || continue
    n=${v#variants/libiactrace_b200_}; run $w $PWD/$v ${n%.so}
  done
done | tee gpurun_out/ab.log
