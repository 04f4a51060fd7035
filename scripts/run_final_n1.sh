#!/bin/bash
# usage (on the GPU box): bash scripts/run_final_n1.sh  -> gpurun_out/final_*
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12) > gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
timeout 200 python scripts/kernel_bench.py 4 576 2304 > gpurun_out/final_kbench.jsonl 2> gpurun_out/final_kbench.err
timeout 100 python scripts/stem_tc_bench.py 576 > gpurun_out/final_stem_tc.jsonl 2> gpurun_out/final_stem_tc.err
timeout 100 python scripts/e2e_timeline.py > gpurun_out/final_timeline.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-range --no-clocks > gpurun_out/final_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hpb_stem_tc -c 1 -o gpurun_out/final_stem_tc python scripts/stem_tc_bench.py 576 > gpurun_out/final_stem_ncu.log 2>&1
tail -4 gpurun_out/final_tests.log; cat gpurun_out/final_smoke.log | tail -2; cut -c1-250 gpurun_out/final_bench_n1.json; cut -c1-200 gpurun_out/final_bench_reference.json; cat gpurun_out/final_stem_tc.jsonl | cut -c1-250
