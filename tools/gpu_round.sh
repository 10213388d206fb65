python bench.py 2>gpurun_out/r1v_bench.err | tail -1 > gpurun_out/r1v_bench.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r1v_bench.json').read())
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['e2e']['phases_of_a_synchronised_pass'])
PY
