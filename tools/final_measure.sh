#!/bin/bash
# One GPU box, one B200: the measurements the round's profiles/ are made from. Writes everything under gpurun_out/.
#   bash tools/final_measure.sh [tag] [kernel regex of the full capture] [matching launches per step]
tag=${1:-r2}
kernels=${2:-k_hist|k_scatter_planes|k_dedup|k_pileup_main|k_median}
per_step=${3:-6}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/gputest_$tag.log 2>&1; tail -2 gpurun_out/gputest_$tag.log
python bench.py --steps 20 --warmup 5 --verify > gpurun_out/${tag}_bench_c2_1gpu.json 2> gpurun_out/bench_err_$tag.log; echo "bench rc=$?"
# launch list of three steps (time + DRAM bytes per kernel); bench values never come from a run under ncu
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${tag}_launches_c2_20M.csv python tools/stage_sweep.py --steps 1 --warmup 2 > gpurun_out/ncu_launch_$tag.log 2>&1; echo "launch list rc=$?"
# full capture of the large kernels of the last step (k_hist, k_scatter_planes, k_dedup x 2, k_pileup_main, k_median)
ncu --set full --clock-control none --import-source on -k "regex:$kernels" \
    --launch-skip $((2 * per_step)) --launch-count $per_step -f -o gpurun_out/prof_$tag python tools/stage_sweep.py --steps 1 --warmup 2 > gpurun_out/ncu_full_$tag.log 2>&1; echo "full capture rc=$?"
