# final single-GPU evidence of a round: tests, smoke, bench (+ reference arm), launch list, ncu captures, side tools
T=${1:-r03z}
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${T}_tests.txt; cat gpurun_out/${T}_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-200 gpurun_out/${T}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${T}.csv python bench.py --steps 2 --warmup 3 > gpurun_out/${T}_under_ncu.log 2>&1
PHX_BENCH_SKIP_DENSE=1 PHX_BENCH_SKIP_EXTRAS=1 ncu --set full --clock-control none --import-source on -k regex:phx_rows -s 6 -c 2 -o gpurun_out/prof_${T} -f python bench.py --steps 2 --warmup 3 > gpurun_out/${T}_under_ncu2.log 2>&1
for c in breast yeast sim690 sim350; do python tools/train_epoch.py --config $c --epochs 3 --many 2>/dev/null | tail -1; done > gpurun_out/${T}_epochs.txt; cut -c1-60,330-420 gpurun_out/${T}_epochs.txt
python tools/sweep_c5.py --cpu-rows 128 2>&1 | grep -E "n_gpus|cpu" > gpurun_out/${T}_sweep_c5_n1.txt; cut -c90-260 gpurun_out/${T}_sweep_c5_n1.txt
python tools/gene_influence.py --count 32 2>&1 | tail -1 > gpurun_out/${T}_influence.txt; cut -c100-300 gpurun_out/${T}_influence.txt
python tools/tc_check.py --shapes 3551,120,1024 11165,200,10000 20000,200,4096 --modes 3xtf32 tf32 --vjp --reps 5 > gpurun_out/${T}_tc_check.txt 2>&1
bash tools/gpu_profile_tc.sh ${T} > /dev/null 2>&1
