#!/bin/bash
for cfg in "ICADV_TC_PERSIST=0" "ICADV_TC_PERSIST=0 ICADV_TC_ONE_CTA=1"; do
  echo "=== $cfg"
  env $cfg timeout 120 python scripts/tile_timeline.py 2>&1 | grep -A1 "conv linear\|conv + GDN\|dgrad" | cut -c1-420
done
