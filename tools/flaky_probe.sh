# Repeats the early part of the GPU suite (the intermittent golden-step failure showed there), then the whole suite.
K='test_diffuse or test_project or fluid_steps or golden_steps'
f=0
for i in $(seq 14); do
  timeout 120 python -m pytest tests -m gpu -x -q -k "$K" > gpurun_out/flaky_$i.log 2>&1
  if grep -q failed gpurun_out/flaky_$i.log; then f=$((f+1)); grep -h "AssertionError\|^FAILED" gpurun_out/flaky_$i.log | cut -c1-400; fi
done
echo "subset: $f failures of 14"
for i in 1 2; do timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -1; done
