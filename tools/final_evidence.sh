#!/bin/bash
# Round-end evidence on one B200 (run under gpurun from the repo root): GPU tests, the bench line, the ncu launch list of
# one training step, the per-shape GEMM table, the txt2img bench.  Every leg has its own timeout; outputs under gpurun_out/.
T=${1:-r02d}
O=gpurun_out
timeout 420 python -m pytest tests -x -q -m gpu > $O/${T}_pytest_gpu.log 2>&1; tail -3 $O/${T}_pytest_gpu.log
timeout 420 python bench.py --steps 30 --warmup 5 > $O/${T}_bench_1gpu.json 2> $O/${T}_bench_1gpu.err; cut -c1-400 $O/${T}_bench_1gpu.json
timeout 480 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --profile-from-start off --csv --log-file $O/${T}_step_metrics_ncu.csv python tools/profile_step.py full > $O/${T}_ncu_step.log 2>&1; tail -2 $O/${T}_ncu_step.log
timeout 200 python tools/gemm_breakdown.py > $O/${T}_gemm_breakdown.log 2>&1; head -2 $O/${T}_gemm_breakdown.log; cp $O/gemm_breakdown.jsonl $O/${T}_gemm_breakdown.jsonl 2>/dev/null
timeout 240 python bench.py --workload txt2img --steps 1 --warmup 1 > $O/${T}_bench_txt2img.json 2> $O/${T}_bench_txt2img.err; cut -c1-300 $O/${T}_bench_txt2img.json
