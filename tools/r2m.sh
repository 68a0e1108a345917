timeout 600 python tools/cold_probe.py 2>&1 | grep -v Warn | cut -c1-200
