#!/bin/bash
# bench.py --gpus 2 with the exchange/gate overlap of csrc/sharded.cu off, on, and its tuning knobs (run under gpurun --gpus 2)
run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NG:-2} --steps 5 --warmup 3 "$@" > gpurun_out/${PFX:-ovl}_$tag.json 2> gpurun_out/${PFX:-ovl}_$tag.err; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/${PFX:-ovl}_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(d["ms_per_step"],2), d["parity"]["ok"], {k:(round(v,2) if isinstance(v,float) else v) for k,v in d["swap"].items() if k!="note"})
except Exception as e: print("$tag", "FAILED", e)
P
}
for v in "$@"; do
  if [ "$v" = off ]; then run off --tune overlap=0; elif [ "$v" = on ]; then run on; else run "on_$v" $(for kv in ${v//,/ }; do echo --tune $kv; done); fi
done
