#!/usr/bin/env bash
# one B200, round-2 measurement batch (everything lands in gpurun_out/)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 300 python bench.py --config cfg5 --steps 3 --warmup 1 > $O/r02_bench_cfg5.json 2> $O/r02_bench_cfg5.err; tail -c 1800 $O/r02_bench_cfg5.json; tail -2 $O/r02_bench_cfg5.err
timeout 120 python scripts/cfg1_latency.py > $O/r02_cfg1_latency.json 2>&1; cat $O/r02_cfg1_latency.json
timeout 300 python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_bench_cfg3_1gpu.json 2> $O/r02_bench_cfg3_1gpu.err; head -c 400 $O/r02_bench_cfg3_1gpu.json; echo
timeout 400 python bench.py --config cfg4 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02_bench_cfg4_1gpu.json 2> $O/r02_bench_cfg4_1gpu.err; head -c 400 $O/r02_bench_cfg4_1gpu.json; echo
# per-kernel table of one cfg4 step (north_star: HBM GB/s and tensor-pipe % per kernel)
PROGEN_PROFILE_RANGE=1 PROGEN_BENCH_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none --profile-from-start off --csv --log-file $O/r02_cfg4_kernels.csv python bench.py --config cfg4 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_cfg4.log 2>&1; tail -1 $O/ncu_cfg4.log | cut -c1-200
# --set full captures of the kernels bench.py's roofline object times (config 2)
ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_fwd_ts|attn_bwd_d|gemm_tc2" -o $O/r02_step_kernels python scripts/dominant_gemm.py > $O/ncu_dom.log 2>&1; tail -2 $O/ncu_dom.log | cut -c1-300
ls -la $O/*.ncu-rep | tail -3
