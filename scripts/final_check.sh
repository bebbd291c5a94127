python -m pytest tests -m gpu -x -q 2>&1 | tail -1
python bench.py > gpurun_out/bench_r1i.json 2> gpurun_out/bench_r1i.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r1i.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"])
print(d["e2e"]["value"], d["e2e"]["ms_per_step"])
print(d["secondary"]["value"], d["secondary"]["ms_per_step"])
print(d["secondary_highdim"]["value"], d["secondary_highdim"]["ms_per_step"])
print(d["secondary_fps"]["ms_per_step"], d["secondary_ball_query"]["ms_per_step"])
PY
