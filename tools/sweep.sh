rm -f gpurun_out/sw_*.jsonl
for b in 0 1 2 3; do
  timeout 100 python tools/microbench.py --n 30 --gs 4 --reps 5 --tune block=$b --out gpurun_out/sw_b$b.jsonl > /dev/null 2>>gpurun_out/mb.err
done
