R() { n=$1; shift; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 "$@"; }
timeout 200 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_multi_1.json 2> gpurun_out/bm1.err; echo rc1=$?
for n in 2 4 8; do R $n bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/bench_multi_$n.json 2> gpurun_out/bm$n.err; echo rc$n=$?; done
R 8 tools/check_multi.py --cfl 300 --visc 0 --steps 4 --h 3072 > gpurun_out/cm8.log 2>&1; echo rccm=$?
R 8 bench.py --gpus 8 --steps 30 --warmup 5 --size 16384 --rows 16384 > gpurun_out/bench_multi_8_16384sq.json 2> gpurun_out/bm8s.err; echo rcs=$?
grep -h "PARITY\|MISMATCH" gpurun_out/cm8.log | tail -3
for f in 1 2 4 8 8_16384sq; do python - <<PY
import json
l=[x for x in open("gpurun_out/bench_multi_$f.json") if x.startswith("{")]
if l:
    j=json.loads(l[-1]); print("$f", j["n_gpus"], round(j["value"]/1e9,2), round(j["ms_per_step"],4), j.get("per_rank_ms_per_step_gpu_and_host_enqueue"))
PY
done
