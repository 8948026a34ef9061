mkdir -p gpurun_out
python tools/ab_elem.py --orders 3 --libs cb2,m3b,cb2m3b --reps 15
python tools/ab_elem.py --orders 2 --libs cb2k2 --reps 15
python tools/ab_elem.py --orders 4 --libs k4cb2m2,k4m2 --reps 15
