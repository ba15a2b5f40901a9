#!/bin/bash
# Round-end validation on one B200: full GPU suite, smoke(), default bench line; `ncu` as 2nd argument adds the ncu launch list of the bench command
# (summarised by scripts/launch_breakdown.py - several minutes of serialised replays).
mkdir -p gpurun_out
tag=${1:-r02f}
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "conv TF", d["roofline"]["achieved"], d["roofline"]["frac"], "launches", d["gpu_launches"], d["clocks"])
print({k:(round(v.get("value",0)),round(v.get("ms_per_step",0),2)) for k,v in d["other_configs"].items()}, "stft frac", d.get("frontend",{}).get("stft",{}).get("frac"))
PY
if [ "$2" = "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 2 --warmup 3 --no-input-pipeline > gpurun_out/bench_under_ncu_${tag}.log 2>&1; echo "ncu rc=$?"
python scripts/launch_breakdown.py gpurun_out/launches_${tag}.csv gpurun_out/${tag}_bench_default_launch_breakdown.txt | head -12
fi
