#!/bin/bash
# e2e (host-buffer) throughput vs number of H2D/kernel/D2H pipeline chunks
for c in 4 8 16 32 64; do
  LWB_E2E_CHUNKS=$c timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > /tmp/e2e_$c.json
  python - "$c" <<'PY'
import json, sys
c = sys.argv[1]
d = json.load(open(f"/tmp/e2e_{c}.json"))
print("chunks", c, "e2e Msamples/s", round(d["e2e"]["value"], 1))
PY
done
