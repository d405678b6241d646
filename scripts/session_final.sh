#!/bin/bash
# final single-GPU evidence run of the round (under gpurun): tests, smoke, bench lines, ncu captures, sanitizer
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2z_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2z_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_reference_arm.json 2> gpurun_out/r2z_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2z_launches.csv python scripts/one_pass.py > gpurun_out/r2z_onepass.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sga_|merge4" -c 9 -o gpurun_out/r2z_sga_full python scripts/one_pass.py 1 8 192 240 624 1 > gpurun_out/r2z_ncu_full.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/debug_hscan.py 1 2 24 16 48 > gpurun_out/r2z_memcheck.log 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python scripts/debug_hscan.py 1 2 24 16 48 > gpurun_out/r2z_racecheck.log 2>&1
tail -3 gpurun_out/r2z_tests.log; cat gpurun_out/r2z_smoke.log | tail -2; cut -c1-400 gpurun_out/r2z_bench_n1.json; cut -c1-300 gpurun_out/r2z_bench_reference_arm.json
grep -E "ERROR SUMMARY" gpurun_out/r2z_memcheck.log gpurun_out/r2z_racecheck.log; tail -2 gpurun_out/r2z_ncu_full.log
