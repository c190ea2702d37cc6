# Round-2 evidence run (on the GPU box): every ncu pass follows a plain run of the same command that exited 0.
set -x
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_bench.log 2>&1
python tools/prof_run.py all 64 > gpurun_out/r02_plain_all.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 6 -c 3 -f -o gpurun_out/r02_step_full python tools/prof_run.py all 64 > gpurun_out/r02_ncu_all.log 2>&1
python tools/prof_run.py bonded 4096 > gpurun_out/r02_plain_b4096.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name regex:bonded_kernel -s 1 -c 1 -f -o gpurun_out/r02_bonded_b4096 python tools/prof_run.py bonded 4096 > gpurun_out/r02_ncu_b4096.log 2>&1
python tools/prof_run.py nb 32 tile5 > gpurun_out/r02_plain_10k.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 4 -c 1 -f -o gpurun_out/r02_nb_pair_10k python tools/prof_run.py nb 32 tile5 > gpurun_out/r02_ncu_10k.log 2>&1
python tools/prof_run.py nb 8 tile24 > gpurun_out/r02_plain_51k.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 4 -c 1 -f -o gpurun_out/r02_nb_block_51k python tools/prof_run.py nb 8 tile24 > gpurun_out/r02_ncu_51k.log 2>&1
tools/micro/issue_bench 3 > gpurun_out/r02_issue_bench.log 2>&1; tools/micro/issue_bench 4 >> gpurun_out/r02_issue_bench.log 2>&1
tools/micro/operand_bench 3 > gpurun_out/r02_operand_bench.log 2>&1; tools/micro/operand_bench 4 >> gpurun_out/r02_operand_bench.log 2>&1
tools/micro/ffma2_bench > gpurun_out/r02_ffma2_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep
