TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29551 tools/check_peer_allreduce.py 2>&1 | tail -2 > gpurun_out/r02u_peer_n8.txt; cat gpurun_out/r02u_peer_n8.txt
for nv in 0 1; do PHX_PEER_NVLS=$nv PHX_BENCH_SKIP_DENSE=1 PHX_BENCH_SKIP_EXTRAS=1 timeout 400 $TR --master-port 2956$nv bench.py --gpus 8 --steps 10 --warmup 3 2>gpurun_out/r02u_bench_n8_nvls$nv.err | tail -1 > gpurun_out/r02u_bench_n8_nvls$nv.json; cut -c1-230 gpurun_out/r02u_bench_n8_nvls$nv.json; done
PHX_BENCH_NCCL=1 PHX_BENCH_SKIP_DENSE=1 PHX_BENCH_SKIP_EXTRAS=1 timeout 400 $TR --master-port 29563 bench.py --gpus 8 --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02u_bench_n8_nccl.json; cut -c1-230 gpurun_out/r02u_bench_n8_nccl.json
timeout 300 $TR --master-port 29571 tools/train_epoch.py --config breast --epochs 3 --many 2>&1 | tail -1 > gpurun_out/r02u_epoch_n8.txt; cat gpurun_out/r02u_epoch_n8.txt
timeout 300 $TR --master-port 29572 tools/sweep_c5.py --strong --global-norm 2>&1 | grep n_gpus > gpurun_out/r02u_sweep_c5_n8_strong.txt; cat gpurun_out/r02u_sweep_c5_n8_strong.txt
timeout 300 $TR --master-port 29573 tools/gene_influence.py --count 16 2>&1 | tail -1 > gpurun_out/r02u_influence_n8.txt; cat gpurun_out/r02u_influence_n8.txt
