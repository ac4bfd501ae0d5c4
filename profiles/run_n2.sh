p=29520
for c in c1 c3 c4; do
  p=$((p+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $p bench.py --gpus 2 --steps 4 --warmup 3 --config $c 2>gpurun_out/${c}_n2.err | grep "^{" > gpurun_out/${c}_n2.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${c}_n2.json").read().strip().splitlines()[-1])
    print("$c", d["ms_per_step"], d["value"], d["e2e"]["value"], d["dp_check"]["ok"])
except Exception as e:
    print("$c failed", e); print(open("gpurun_out/${c}_n2.err").read()[-600:])
PY
done
