python tools/ab_elem.py --orders 3,2,4 --libs base --reps 15
timeout 400 python -m pytest tests -m gpu -x -q -k "assembly or driver or permuted or delaunay or lu_path or tau" 2>&1 | tail -3
