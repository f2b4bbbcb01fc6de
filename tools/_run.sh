nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_multirank.py -q -s -m gpu -p no:cacheprovider > gpurun_out/mr8.log 2>&1; tail -4 gpurun_out/mr8.log
for n in 8 4 2; do
for mode in auto nccl; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --gather $mode > gpurun_out/bench${n}_$mode.json 2> gpurun_out/bench${n}_$mode.err
echo "N=$n $mode exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench${n}_$mode.json')); print(' weak', d['value'], 'e2e', d['e2e']['value']); [print('  strong', s['global_batch'], s['value'], s['ms_per_step'], s['bit_identical_to_single_gpu'], s['gather'], s['global_checksum']) for s in d['strong']]"
done; done
