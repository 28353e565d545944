#!/bin/bash
N=$(python -c "import torch; print(torch.cuda.device_count())")
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_multirank.py -q -m gpu --timeout 500 > gpurun_out/r02_mgpu_tests_n$N.log 2>&1
echo "mgpu tests exit $?" | tee -a gpurun_out/r02_mgpu_tests_n$N.log; tail -2 gpurun_out/r02_mgpu_tests_n$N.log
timeout 400 $TR --master-port 29611 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err
echo "bench rc $?"; tail -c 1500 gpurun_out/r02_bench_n${N}.json
timeout 300 $TR --master-port 29612 tools/exchange_bench.py > gpurun_out/r02_exchange_n${N}.json 2> gpurun_out/r02_exchange_n${N}.err
echo "exchange rc $?"; cat gpurun_out/r02_exchange_n${N}.json
timeout 400 $TR --master-port 29613 tools/fed_round_bench.py --preset c3 --rounds 4 > gpurun_out/r02_fed_c3_n${N}.json 2> gpurun_out/r02_fed_c3_n${N}.err
echo "c3 rc $?"; cat gpurun_out/r02_fed_c3_n${N}.json
timeout 400 $TR --master-port 29614 tools/fed_round_bench.py --preset c4 --rounds 3 > gpurun_out/r02_fed_c4_n${N}.json 2> gpurun_out/r02_fed_c4_n${N}.err
echo "c4 rc $?"; cat gpurun_out/r02_fed_c4_n${N}.json
timeout 300 $TR --master-port 29615 tools/infer_bench.py > gpurun_out/r02_infer_c5_n${N}.json 2> gpurun_out/r02_infer_c5_n${N}.err
echo "c5 rc $?"; cat gpurun_out/r02_infer_c5_n${N}.json
