N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 3 --warmup 3 2> gpurun_out/bench_power_scan_n$N.err | grep '^{' > gpurun_out/bench_power_scan_n$N.json
python -c "
import json; d=json.load(open('gpurun_out/bench_power_scan_n$N.json')); print($N, d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['strong']['value'])"
