#!/bin/bash
# multi-GPU call: multirank tests + bench at N = all visible GPUs (+ sharded transport timing)
N=$(python -c "import torch; print(torch.cuda.device_count())")
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multirank.py -q -m gpu --timeout 800 > gpurun_out/r02_mgpu_tests_n$N.log 2>&1
echo "exit $?" >> gpurun_out/r02_mgpu_tests_n$N.log
tail -3 gpurun_out/r02_mgpu_tests_n$N.log
for tr in auto p2p_sharded; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
  bench.py --gpus $N --steps 30 --warmup 5 --fed-transport $tr > gpurun_out/r02_bench_n${N}_$tr.json 2> gpurun_out/r02_bench_n${N}_$tr.err
echo "bench $tr rc $?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n${N}_$tr.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ["value","ms_per_step","fedavg_exchange_ms","fedavg_transport","fedavg_bitexact","fedavg_exchange_nvlink_gbs_per_rank","fedavg_round_s","fed_round_c4","clocks"]})
PY
tail -3 gpurun_out/r02_bench_n${N}_$tr.err
done
