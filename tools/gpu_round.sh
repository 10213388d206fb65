python tools/plan_probe.py 2>&1 | grep -v Warn | tail -4
