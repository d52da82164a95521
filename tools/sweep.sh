rm -f gpurun_out/sw_*.jsonl
timeout 100 python tools/microbench.py --n 30 --gs 4 --reps 5 --out gpurun_out/sw_auto.jsonl > /dev/null 2>>gpurun_out/mb.err
timeout 100 python tools/microbench.py --n 30 --gs 4 --reps 5 --tune tile=2 --out gpurun_out/sw_tile.jsonl > /dev/null 2>>gpurun_out/mb.err
