#!/bin/bash
# development: e2e of bench.py for several head-chunk sizes of the host-staged pipeline
for hc in 8 16 32; do
  python bench.py --no-other-paths --no-cpu-baseline --steps 3 --e2e-heads-per-chunk $hc 2>/dev/null > /tmp/e2e_$hc.json
  python - "$hc" <<'PY'
import json, sys
d = json.load(open("/tmp/e2e_%s.json" % sys.argv[1]))
print("hc", sys.argv[1], "e2e", round(d["e2e"]["value"], 1), "TOPS  pcie", round(d["pcie"]["duplex_GBs_each_way_per_gpu"], 1), "GB/s")
PY
done
