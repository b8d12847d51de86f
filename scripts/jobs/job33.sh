ACE_DIAG_CLUSTER=8 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NC8', d['ms_per_step'], d['roofline']['phase_ms'])"
ACE_DIAG_CLUSTER=16 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NC16', d['ms_per_step'], d['roofline']['phase_ms'])"
