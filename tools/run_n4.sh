TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29651 tools/check_peer_allreduce.py 2>&1 | tail -1 > gpurun_out/r03g_peer_n4.txt; cat gpurun_out/r03g_peer_n4.txt
PHX_BENCH_SKIP_DENSE=1 PHX_BENCH_SKIP_EXTRAS=1 timeout 400 $TR --master-port 29652 bench.py --gpus 4 --steps 10 --warmup 3 2>gpurun_out/r03g_bench_n4.err | tail -1 > gpurun_out/r03g_bench_n4.json; cut -c1-230 gpurun_out/r03g_bench_n4.json; tail -2 gpurun_out/r03g_bench_n4.err
timeout 300 $TR --master-port 29653 tools/train_epoch.py --config breast --epochs 3 --many 2>&1 | tail -1 > gpurun_out/r03g_epoch_n4.txt; cut -c1-600 gpurun_out/r03g_epoch_n4.txt
timeout 300 $TR --master-port 29654 tools/sweep_c5.py --strong --global-norm 2>&1 | grep n_gpus > gpurun_out/r03g_sweep_c5_n4_strong.txt; cut -c90-260 gpurun_out/r03g_sweep_c5_n4_strong.txt
