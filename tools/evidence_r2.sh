# One-GPU evidence pass: tests, every bench configuration, ncu launch list and full-set capture of the same command.
# usage (on a GPU box): bash tools/evidence_r2.sh TAG   -> gpurun_out/ev_TAG_*
T=${1:-r2}; O=gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) > $O/ev_${T}_tests.log; echo tests rc=$?
timeout 400 python bench.py > $O/ev_${T}_bench_headline.json 2> $O/ev_${T}_bench_headline.err; echo headline rc=$?
for c in scene250 scene128 inject1024 ksweep; do
  timeout 400 python bench.py --config $c > $O/ev_${T}_bench_$c.json 2> $O/ev_${T}_bench_$c.err; echo $c rc=$?
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/ev_${T}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-variants --sustain-seconds 0 > $O/ev_${T}_ncu_launch.log 2>&1; echo launches rc=$?
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_jacobi_tile|k_divergence4|k_gradient4|k_advect" -s 96 -c 16 \
  -f -o $O/ev_${T}_full python bench.py --steps 2 --warmup 3 --no-cpu --no-variants --sustain-seconds 0 > $O/ev_${T}_ncu_full.log 2>&1; echo full rc=$?
ncu -i $O/ev_${T}_full.ncu-rep --page raw --csv > $O/ev_${T}_full.raw.csv 2>/dev/null
