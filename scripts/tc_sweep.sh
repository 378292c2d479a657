#!/bin/bash
# developer sweep: per-CTA timelines of the contraction kernel under different ring depths / CTA residency
for cfg in "ICADV_TC_S=0" "ICADV_TC_ONE_CTA=1"; do
  echo "=== $cfg"
  env $cfg timeout 120 python scripts/tile_timeline.py 2>&1 | cut -c1-330
done
