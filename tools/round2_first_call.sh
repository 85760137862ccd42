#!/bin/bash
# First GPU call of the next round, in one gpurun (everything lands in gpurun_out/):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# 1. the whole GPU test suite (the experimental variants report XPASS / XFAIL, they cannot fail it)
# 2. A/B of every tile-kernel variant (ie_set_option("encode_variant", 0..7)) with parity, torch-free
# 3. the other SURVEY 8 rows with the default kernels and with decode_variant=1
# 4. bench.py (N=1), then the ncu launch list of the same command
# Each step has its own timeout so that a hang in an experimental kernel cannot take the box with it.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rxX > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/r2_steps.log
timeout 300 python tools/ab_quick.py > gpurun_out/r2_ab_quick.log 2>&1; echo "ab_quick rc=$?" | tee -a gpurun_out/r2_steps.log
timeout 300 python tools/ab_quick_rows.py > gpurun_out/r2_rows_default.log 2>&1; echo "rows(default) rc=$?" | tee -a gpurun_out/r2_steps.log
cp gpurun_out/ab_quick_rows.json gpurun_out/r2_rows_default.json 2>/dev/null
timeout 300 python tools/ab_quick_rows.py --options decode_variant=1 > gpurun_out/r2_rows_dec1.log 2>&1; echo "rows(decode_variant=1) rc=$?" | tee -a gpurun_out/r2_steps.log
cp gpurun_out/ab_quick_rows.json gpurun_out/r2_rows_dec1.json 2>/dev/null
timeout 300 python tools/ab_quick_rows.py --options me_variant=1 > gpurun_out/r2_rows_me1.log 2>&1; echo "rows(me_variant=1) rc=$?" | tee -a gpurun_out/r2_steps.log
cp gpurun_out/ab_quick_rows.json gpurun_out/r2_rows_me1.json 2>/dev/null
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?" | tee -a gpurun_out/r2_steps.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1; echo "ncu launches rc=$?" | tee -a gpurun_out/r2_steps.log
tail -3 gpurun_out/r2_pytest_gpu.log; grep -h "ms/encode" gpurun_out/r2_ab_quick.log | tail -12; cat gpurun_out/r2_bench.json | cut -c1-600
