# 8-GPU evidence of a round (one box): peer all-reduce check, default bench (extras included), epoch, config-5 sweep strong
# and weak, gene-influence scan
T=${1:-r03n8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29751 tools/check_peer_allreduce.py 2>&1 | tail -1 > gpurun_out/${T}_peer.txt; cat gpurun_out/${T}_peer.txt
timeout 900 $TR --master-port 29752 bench.py --gpus 8 --steps 20 --warmup 3 2>gpurun_out/${T}_bench.err | tail -1 > gpurun_out/${T}_bench.json; cut -c1-230 gpurun_out/${T}_bench.json
timeout 300 $TR --master-port 29753 tools/train_epoch.py --config breast --epochs 3 --many 2>&1 | tail -1 > gpurun_out/${T}_epoch.txt; cut -c330-420 gpurun_out/${T}_epoch.txt
timeout 300 $TR --master-port 29754 tools/sweep_c5.py --strong --global-norm 2>&1 | grep n_gpus > gpurun_out/${T}_sweep_c5_strong.txt; cut -c90-260 gpurun_out/${T}_sweep_c5_strong.txt
timeout 300 $TR --master-port 29755 tools/sweep_c5.py 2>&1 | grep n_gpus > gpurun_out/${T}_sweep_c5_weak.txt; cut -c90-260 gpurun_out/${T}_sweep_c5_weak.txt
timeout 300 $TR --master-port 29756 tools/gene_influence.py --count 32 2>&1 | tail -1 > gpurun_out/${T}_influence.txt; cut -c100-300 gpurun_out/${T}_influence.txt
