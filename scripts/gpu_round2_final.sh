#!/usr/bin/env bash
# one B200: the round-2 end-of-round batch (everything lands in gpurun_out/)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15) > $O/r02_final_tests.log; cat $O/r02_final_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 > $O/r02_bench_final.json 2> $O/r02_bench_final.err; tail -c 400 $O/r02_bench_final.json; tail -2 $O/r02_bench_final.err
timeout 300 python bench.py --config cfg5 --steps 3 --warmup 3 > $O/r02_bench_cfg5.json 2> $O/r02_bench_cfg5.err; tail -c 1500 $O/r02_bench_cfg5.json; tail -2 $O/r02_bench_cfg5.err
(timeout 200 python scripts/decode_profile.py 1 64) > $O/r02_decode_profile_final.json 2> $O/r02_decode_profile_final.err
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
