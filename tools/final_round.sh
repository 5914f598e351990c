set -u
mkdir -p gpurun_out
bash tools/profile_round.sh r04 > gpurun_out/r04_profile.log 2>&1
for c in 2 3 4 5; do
  python bench.py --config $c > gpurun_out/r04_cfg${c}.json 2> gpurun_out/r04_cfg${c}.err
done
python __graft_entry__.py smoke > gpurun_out/r04_smoke.log 2>&1
tail -2 gpurun_out/r04_smoke.log
for c in 2 3 4 5; do python - <<PY
import json
d=json.load(open("gpurun_out/r04_cfg${c}.json"))
print(${c}, round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "gemm frac", round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz"])
PY
done
