timeout 600 python scripts/concurrent_fits.py C5 2>&1 | tail -5
timeout 600 python scripts/concurrent_fits.py C2 2>&1 | tail -5
