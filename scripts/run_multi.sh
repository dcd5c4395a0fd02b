#!/bin/bash
# bench.py on N GPUs of one box (the driver's launch line), result -> gpurun_out/bench_nN.json
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --verify > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -1 gpurun_out/bench_n$N.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('Mrays/s', round(d['value'], 1), 'ms/step', round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1),
      'identical', d.get('multi_gpu_frame_identical_to_1gpu'), 'accel', round(d['accelerated_mode']['value'], 1))
print(d['per_rank']['rows'])"
