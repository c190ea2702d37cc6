set -x
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_bench.log 2>&1
python tools/prof_run.py all 64 > gpurun_out/r02_plain_all.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 6 -c 3 -f -o gpurun_out/r02_step_full python tools/prof_run.py all 64 > gpurun_out/r02_ncu_all.log 2>&1
python tools/prof_run.py bonded 4096 > gpurun_out/r02_plain_b4096.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 2 -c 1 -f -o gpurun_out/r02_bonded_b4096 python tools/prof_run.py bonded 4096 > gpurun_out/r02_ncu_b4096.log 2>&1
python tools/prof_run.py nb 128 tile5 > gpurun_out/r02_plain_blk.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 4 -c 2 -f -o gpurun_out/r02_nb_block_10k python tools/prof_run.py nb 128 tile5 > gpurun_out/r02_ncu_blk.log 2>&1
MLP_NB_BLOCKS=0 python tools/prof_run.py nb 128 tile5 > gpurun_out/r02_plain_blk0.log 2>&1 && \
MLP_NB_BLOCKS=0 ncu --set full --clock-control none --import-source on -s 4 -c 1 -f -o gpurun_out/r02_nb_pair_10k python tools/prof_run.py nb 128 tile5 > gpurun_out/r02_ncu_blk0.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail; tail -2 gpurun_out/r02_ncu_all.log gpurun_out/r02_ncu_blk.log
