#!/bin/bash
# First GPU call of the next round: the two candidates staged behind compile-time flags, each as a full
# variant library next to the default build.  Run HERE first:   bash tools/build_full_variant.sh zero "-DHDG_ZERO_ASYNC=1"
#                                                               bash tools/build_full_variant.sh gen  "-DHDG_MG_GENERAL=1"
# then:   gpurun --timeout 600 -- 'bash tools/round2_first_call.sh'
mkdir -p gpurun_out
V=$PWD/hdiscontinuousgalerkin.jl_b200/variants
echo "== default build"; timeout 100 python bench.py --steps 50 --no-cpu --no-pcg --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4e ms/step %.4f kernel_ms %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
if [ -f $V/lib_zero.so ]; then
  echo "== HDG_ZERO_ASYNC: full GPU suite, then the same bench"
  HDG_B200_LIB=$V/lib_zero.so timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
  HDG_B200_LIB=$V/lib_zero.so timeout 100 python bench.py --steps 50 --no-cpu --no-pcg --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.4e ms/step %.4f kernel_ms %.4f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
fi
if [ -f $V/lib_gen.so ]; then
  echo "== HDG_MG_GENERAL: multigrid tests incl. test_multigrid_term_on_unstructured_meshes"
  HDG_B200_LIB=$V/lib_gen.so timeout 200 python -m pytest tests -m gpu -x -q -k "multigrid" 2>&1 | tail -4
fi
