python tools/ab_elem.py --orders 3,2,4 --libs rtg --reps 15
