set -x
P=29611
for N in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_train_dp$N.json 2> gpurun_out/dp$N.err; echo rc=$?; tail -c 200 gpurun_out/dp$N.err
P=$((P+1))
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29620 bench.py --gpus 8 --steps 20 --warmup 5 --no-overlap > gpurun_out/r2_bench_train_dp8_no_overlap.json 2> gpurun_out/dp8b.err; echo rc=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 20 --warmup 5 --config config/fern_batch_h256.json > gpurun_out/r2_bench_fern_pinhole_dp8.json 2> gpurun_out/dp8c.err; echo rc=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 8 --steps 20 --warmup 5 --config config/fern_batch_h256.json --rays ndc > gpurun_out/r2_bench_fern_ndc_dp8.json 2> gpurun_out/dp8d.err; echo rc=$?
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29623 inference.py --config config/lego_batch_h256.json --frames 40 --out gpurun_out/frames8.npy > gpurun_out/r2_inference_8gpu.log 2>&1; echo rc=$?; tail -2 gpurun_out/r2_inference_8gpu.log; rm -f gpurun_out/frames8.npy
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -2
