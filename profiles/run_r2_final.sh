#!/bin/bash
# round 2, final state: bench lines, launch list of the bench command, ncu --set full of the headline kernel, all configs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/final
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/final/bench_1gpu_c4.json 2> gpurun_out/final/bench_c4.err; tail -c 200 gpurun_out/final/bench_c4.err
timeout 900 python bench.py --config c5 --steps 10 --warmup 3 > gpurun_out/final/bench_1gpu_c5.json 2> gpurun_out/final/bench_c5.err; tail -c 200 gpurun_out/final/bench_c5.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_reference_arm.json 2> gpurun_out/final/bench_ref.err; tail -c 200 gpurun_out/final/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/final/ncu_launches.log 2>&1
tail -2 gpurun_out/final/ncu_launches.log | cut -c1-200
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'eval_tile_kernel' -c 2 -f -o gpurun_out/final/tile_final python profiles/prof_r02.py > gpurun_out/final/ncu_tile.log 2>&1
tail -3 gpurun_out/final/ncu_tile.log | cut -c1-200
timeout 1500 python profiles/bench_configs.py > gpurun_out/final/configs.jsonl 2> gpurun_out/final/configs.err; tail -c 300 gpurun_out/final/configs.err; wc -l gpurun_out/final/configs.jsonl
