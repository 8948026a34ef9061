python tools/ab_elem.py --orders 4 --libs base,k4m4 --reps 15
python tools/ab_elem.py --orders 3,2 --libs base --reps 10
