"""Runs tools/order_probe.py the way tests/test_ordering_gpu.py does (subprocess, captured pipes) and stores the output."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = sys.argv[1]
env = dict(os.environ, FRUID_B200_LIB="debug")
r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "order_probe.py"), "--quick", "--assert-safe"] + sys.argv[2:], env=env,
                   capture_output=True, text=True, timeout=900)
open(out, "w").write(r.stdout + "\n--- stderr ---\n" + r.stderr + f"\nrc={r.returncode}\n")
print(out, "rc", r.returncode)
