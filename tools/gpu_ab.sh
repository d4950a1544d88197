#!/bin/bash
# one gpurun visit for the flat-epilogue / bias-K-step chain kernels: A/B parity + timing against the previous
# kernels, the GPU test suite, the bench line, the tensor-pipe metrics of the new kernels
mkdir -p gpurun_out
V=$PWD/robust-e-nerf_b200/ren_b200/lib/variants
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/gpu.txt 2>&1
timeout 200 python tools/ab_kernels.py new > gpurun_out/ab_new.json 2> gpurun_out/ab_new.err; echo "ab_new rc=$?"
REN_B200_LIB=$V/libren_b200_old.so timeout 120 python tools/ab_kernels.py old > gpurun_out/ab_old.json 2> gpurun_out/ab_old.err; echo "ab_old rc=$?"
timeout 420 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 150 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:"ren_tc_chain|ren_tc_wgrad" -s 12 -c 3 --csv --log-file gpurun_out/ncu_chain.csv \
  python bench.py --steps 1 --warmup 3 --min-ms 0 --graph off --no-secondary --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
REN_B200_LIB=$V/libren_b200_flat.so timeout 120 python tools/ab_kernels.py flat > gpurun_out/ab_flat.json 2> gpurun_out/ab_flat.err; echo "ab_flat rc=$?"
cat gpurun_out/ab_new.json gpurun_out/ab_old.json gpurun_out/ab_flat.json; tail -3 gpurun_out/ab_new.err
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cut -c1-600 gpurun_out/bench.log
