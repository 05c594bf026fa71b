# how profiles/r2_bench_n2/4/8*.json were produced: gpurun --gpus 8 -- 'bash tools/run_scaling.sh'
set -x
mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -m gpu -q -x 2>&1 | tail -2 > gpurun_out/r2s_multigpu_test.log; cat gpurun_out/r2s_multigpu_test.log
for n in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n > gpurun_out/r2s_n$n.out 2> gpurun_out/r2s_n$n.err; echo rc=$?; tail -1 gpurun_out/r2s_n$n.out > gpurun_out/r2s_n$n.json
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --workload 64ch_60k_x4 > gpurun_out/r2s_n8_64ch.out 2> gpurun_out/r2s_n8_64ch.err; echo rc=$?; tail -1 gpurun_out/r2s_n8_64ch.out > gpurun_out/r2s_n8_64ch.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s_ref.out 2>&1; tail -1 gpurun_out/r2s_ref.out > gpurun_out/r2s_ref.json
