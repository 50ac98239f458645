# BASELINE.json configs[4] on N GPUs of one box: bash tools/scale16k.sh N  (N = 1, 2, 4, 8) -> gpurun_out/s16k_*_N.json
# weak: 16384 x 2048*N, strong: 16384 x 16384; plus the headline weak-scaling line (4096 x 4096*N) for the same box.
N=${1:-1}; O=gpurun_out
run() { if [ "$N" = 1 ]; then timeout 400 python bench.py --gpus 1 "$@"; else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; fi; }
for c in weak16k strong16k headline; do
  run --config $c --steps 20 --warmup 3 --no-cpu --no-variants --sustain-seconds 0 > $O/s16k_${c}_$N.json 2> $O/s16k_${c}_$N.err; echo $c rc=$?
  python - <<PY
import json
l=[x for x in open("$O/s16k_${c}_$N.json") if x.startswith("{")]
if l:
    j=json.loads(l[-1]); print("$c", j["n_gpus"], j["config"]["grid"], round(j["value"]/1e9,2), "G", round(j["ms_per_step"],4), "ms", j.get("parity_vs_1gpu","")[:40], (j.get("e2e") or {}).get("value"))
PY
done
