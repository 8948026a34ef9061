timeout 500 python -m pytest tests -m gpu -x -q -k "multigrid or mg or driver or size" 2>&1 | tail -4
HDG_MG_TRACE=1 python tools/mg_trace.py 1 1000 500
HDG_MG_NOFUSE=1 HDG_MG_TRACE=1 python tools/mg_trace.py 1 1000 500 | head -2
HDG_MG_TRACE=1 python tools/mg_trace.py 3 2000 1000 | head -2
HDG_MG_NOFUSE=1 HDG_MG_TRACE=1 python tools/mg_trace.py 3 2000 1000 | head -2
