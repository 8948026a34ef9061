mkdir -p gpurun_out
python tools/ab_elem.py --orders 3,2,4 --libs head --reps 15
HDG_MG_TRACE=1 python tools/mg_trace.py 1 1000 500
HDG_MG_TRACE=1 python tools/mg_trace.py 3 2000 1000
