p=29540
for c in c1 c4 c3; do
  p=$((p+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $p bench.py --gpus 8 --steps 4 --warmup 3 --config $c 2>gpurun_out/${c}_n8.err | grep "^{" > gpurun_out/${c}_n8.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${c}_n8.json").read().strip().splitlines()[-1])
    print("$c", d["ms_per_step"], d["value"], d["e2e"]["value"], d["dp_check"]["ok"])
except Exception as e:
    print("$c failed", e); print(open("gpurun_out/${c}_n8.err").read()[-600:])
PY
done
