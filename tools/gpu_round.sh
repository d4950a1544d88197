#!/bin/bash
# one gpurun visit: GPU test suite, smoke, the bench line, tensor-pipe / DRAM metrics of the three big kernels, and the
# ncu launch list of a step (each under its own timeout; outputs in gpurun_out/)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 420 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 150 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:"ren_tc_chain|ren_tc_wgrad" -s 12 -c 3 --csv --log-file gpurun_out/ncu_chain.csv \
  python bench.py --steps 1 --warmup 3 --min-ms 0 --graph off --no-secondary --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu metrics rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 230 -c 170 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --min-ms 0 --graph off --no-secondary --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench.log
