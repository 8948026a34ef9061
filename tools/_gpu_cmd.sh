mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu2.log
timeout 300 python tools/ab_elem.py --orders 2,4 --libs base0,s4 --reps 20 > gpurun_out/ab_stage.txt 2>&1; cat gpurun_out/ab_stage.txt
