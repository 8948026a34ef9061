mkdir -p gpurun_out
for fm in 0 160 640 2304 8192; do
HDG_MG_FUSE_MAX=$fm timeout 200 python bench.py --no-cpu --steps 5 --pcg mg > gpurun_out/b_$fm.json 2> gpurun_out/b_$fm.err; 
python - <<PY
import json
d=json.loads(open('gpurun_out/b_$fm.json').read().strip().splitlines()[-1])
print("fuse_max $fm", d["pcg"]["iterations"], "solve_s %.5f ms/iter %.4f"%(d["pcg"]["solve_s"], d["pcg"]["ms_per_iter"]), "driver", d["e2e"]["driver"].get("ms_steps"))
PY
done
