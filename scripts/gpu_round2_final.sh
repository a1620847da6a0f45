#!/usr/bin/env bash
# one B200: the round-2 end-of-round batch (everything lands in gpurun_out/)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15) > $O/r02_final_tests.log; cat $O/r02_final_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 > $O/r02_bench_final.json 2> $O/r02_bench_final.err; tail -c 700 $O/r02_bench_final.json; tail -2 $O/r02_bench_final.err
timeout 300 python bench.py --config cfg5 --steps 3 --warmup 3 > $O/r02_bench_cfg5.json 2> $O/r02_bench_cfg5.err; tail -c 1500 $O/r02_bench_cfg5.json; tail -2 $O/r02_bench_cfg5.err
timeout 400 python bench.py --config cfg4 --steps 5 --warmup 3 --no-cpu-baseline > $O/r02_bench_cfg4_1gpu.json 2> $O/r02_bench_cfg4_1gpu.err; head -c 300 $O/r02_bench_cfg4_1gpu.json; echo
PROGEN_PROFILE_RANGE=1 PROGEN_BENCH_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none --profile-from-start off --csv --log-file $O/r02_cfg4_kernels.csv python bench.py --config cfg4 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_cfg4.log 2>&1; tail -1 $O/ncu_cfg4.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
