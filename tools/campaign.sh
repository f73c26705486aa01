#!/bin/bash
# Measurement campaign on N GPUs of one box: tools/campaign.sh N TAG [what...]
# Every line lands in gpurun_out/TAG_<name>_n<N>.json (last line of stdout = the bench JSON).
N=$1; TAG=$2; shift 2
WHAT="$@"; [ -z "$WHAT" ] && WHAT="weak strong thermalA thermalB c3 c4"
mkdir -p gpurun_out
run() { # name, bench args...
  name=$1; shift
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 "$@" 2> gpurun_out/${TAG}_${name}_n$N.err | grep '^{' | tail -1 > gpurun_out/${TAG}_${name}_n$N.json
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N "$@" 2> gpurun_out/${TAG}_${name}_n$N.err | grep '^{' | tail -1 > gpurun_out/${TAG}_${name}_n$N.json
  fi
  python - gpurun_out/${TAG}_${name}_n$N.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read())
    print(sys.argv[1], "ms/step %.2f" % d["ms_per_step"], "updates/s %.3e" % d["value"], {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()}, d.get("slow_path", {}).get("wide_trajectories_per_update"), d["checks"])
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
}
for w in $WHAT; do
  case $w in
    weak) run weak --steps 10 --warmup 3 --no-e2e --no-cpu ;;
    strong) [ "$N" != "1" ] && run strong --steps 10 --warmup 3 --no-e2e --no-cpu --no-parity --scaling strong ;;
    thermalA) run thermalA --config thermal --scaling strong --grid 1024 512 1024 --ppc 2 --steps 8 --warmup 3 --no-e2e ;;
    thermalB) [ "$N" != "1" ] && run thermalB --config thermal --scaling strong --grid 1024 1024 1024 --ppc 2 --steps 8 --warmup 3 --no-e2e ;;
    c3) run c3 --config lwfa_like --steps 10 --warmup 3 --no-e2e ;;
    c4) run c4 --config foil_like --steps 6 --warmup 3 --no-e2e ;;
    lwfa) [ "$N" != "1" ] && run lwfa --config lwfa --ppc 8 --steps 900 --warmup 3 --no-e2e ;;
    hot) run hot --config lwfa_hot --shape TSC --grid 256 256 256 --ppc 8 --steps 6 --warmup 3 --no-e2e ;;
    hotcic) run hotcic --config lwfa_hot --grid 256 256 256 --ppc 8 --steps 6 --warmup 3 --no-e2e ;;
  esac
done
