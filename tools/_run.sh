nvidia-smi -L
python -m pytest tests/test_gpu_multirank.py -q -s -m gpu -p no:cacheprovider > gpurun_out/mr_r02d.log 2>&1; tail -15 gpurun_out/mr_r02d.log
for mode in auto nccl; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --gather $mode > gpurun_out/bench2_$mode.json 2> gpurun_out/bench2_$mode.err
echo "exit $?"; tail -3 gpurun_out/bench2_$mode.err; python -c "
import json; d=json.load(open('gpurun_out/bench2_$mode.json')); print(d['value'], d['e2e']['value']); [print(s) for s in d['strong']]"
done
