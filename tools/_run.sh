for ug in 1 2 3; do
python tools/expect_bench.py --n 30 --small --tune expect_ug=$ug >> gpurun_out/s32_expect_n30.jsonl 2>> gpurun_out/s32.err
python tools/expect_bench.py --n 26 --small --reps 50 --tune expect_ug=$ug >> gpurun_out/s32_expect_n26.jsonl 2>> gpurun_out/s32.err
done
cat gpurun_out/s32_expect_n30.jsonl gpurun_out/s32_expect_n26.jsonl | cut -c1-200
python -m pytest tests/test_gates_gpu.py -m gpu -x -q -k expect 2>&1 | tail -2
tail -5 gpurun_out/s32.err
