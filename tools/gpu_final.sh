#!/bin/bash
# final measurements of the round at N = visible GPUs: bench line (+ reference arm at N=1), ncu launch list at N=1
N=$(python -c "import torch; print(torch.cuda.device_count())")
V=${1:-v19}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r02_bench_$V.json 2> gpurun_out/r02_bench_$V.err; echo "bench rc $?"
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_$V.json 2>/dev/null; echo "ref rc $?"
  timeout 300 python tools/profile_step.py > /dev/null 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r02_launches_eager_step_$V.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
  echo "launch list rc $?"
  timeout 300 python tools/infer_bench.py > gpurun_out/r02_infer_c5_n1_$V.json 2>/dev/null; echo "c5 rc $?"
  timeout 400 python tools/fed_round_bench.py --preset c3 --rounds 3 > gpurun_out/r02_fed_c3_n1_$V.json 2>/dev/null; echo "c3 rc $?"
else
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  timeout 600 $TR --master-port 29611 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02_bench_n${N}_$V.json 2> gpurun_out/r02_bench_n${N}_$V.err; echo "bench rc $?"
  timeout 400 $TR --master-port 29614 tools/fed_round_bench.py --preset c4 --rounds 3 > gpurun_out/r02_fed_c4_n${N}_$V.json 2> /dev/null; echo "c4 rc $?"
fi
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_*_$V.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    keys=["value","ms_per_step","e2e","fedavg_exchange_ms","fedavg_transport","fedavg_bitexact","fedavg_round_s","round_images_per_s","local_training_s"]
    print(f, {k:d.get(k) for k in keys if k in d}, (d.get("fed_round_c4") or {}).get("round_s"), (d.get("roofline") or {}).get("frac"))
PY
